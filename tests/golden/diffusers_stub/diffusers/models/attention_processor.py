from oracle import d31


class Attention(d31.Attention):
    """diffusers keyword spelling -> oracle/d31.Attention (self-attention uses only)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False,
                 qk_norm=None, **kw):
        if cross_attention_dim is not None:
            raise NotImplementedError("cross-attention is outside the DiT shim")
        super().__init__(query_dim, heads, dim_head, bias=bias, qk_norm=qk_norm, **kw)
