from . import transformer_temporal  # noqa: F401
