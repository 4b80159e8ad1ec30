from oracle.d31 import TimestepEmbedding, Timesteps  # noqa: F401
