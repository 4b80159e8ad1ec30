"""VaeImageProcessor for tensors: preprocess normalises [0, 1] -> [-1, 1], postprocess
(output_type "pt") denormalises to [0, 1]."""


class VaeImageProcessor:
    def __init__(self, vae_scale_factor=8, **unused):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height=None, width=None):
        """Tensors in [0, 1] -> [-1, 1] (sizes already multiples of the VAE factor)."""
        return 2.0 * image - 1.0

    def postprocess(self, image, output_type="pt"):
        if output_type == "latent":
            return image
        if output_type != "pt":
            raise NotImplementedError(output_type)
        return (image / 2 + 0.5).clamp(0, 1)
