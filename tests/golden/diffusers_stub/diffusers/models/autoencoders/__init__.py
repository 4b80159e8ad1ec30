from . import autoencoder_kl_cogvideox  # noqa: F401
