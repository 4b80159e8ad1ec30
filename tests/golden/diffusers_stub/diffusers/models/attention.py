from oracle.d31 import FeedForward  # noqa: F401
from oracle.unet import BasicTransformerBlock as _BTB


class BasicTransformerBlock(_BTB):
    """diffusers positional spelling (dim, num_attention_heads, attention_head_dim,
    cross_attention_dim=...) as used at crossview_temporal.py:295-298."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, cross_attention_dim=None,
                 **unused):
        super().__init__(dim, num_attention_heads, attention_head_dim,
                         cross_attention_dim=cross_attention_dim)
