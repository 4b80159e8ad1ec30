"""VaeImageProcessor.postprocess for tensors (output_type "pt"): denormalise to [0, 1]."""


class VaeImageProcessor:
    def __init__(self, vae_scale_factor=8, **unused):
        self.vae_scale_factor = vae_scale_factor

    def postprocess(self, image, output_type="pt"):
        if output_type == "latent":
            return image
        if output_type != "pt":
            raise NotImplementedError(output_type)
        return (image / 2 + 0.5).clamp(0, 1)
