from oracle import d31
from oracle.unet import _CrossAttention


def Attention(query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False,
              qk_norm=None, **kw):
    """diffusers keyword spelling -> oracle restatements: self-attention (optionally with
    per-head RMSNorm) or cross-attention."""
    if cross_attention_dim is not None:
        assert qk_norm is None and not bias
        return _CrossAttention(query_dim, cross_attention_dim, heads, dim_head)
    return d31.Attention(query_dim, heads, dim_head, bias=bias, qk_norm=qk_norm, **kw)
