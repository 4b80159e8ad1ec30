import torch


def randn_tensor(shape, generator=None, device=None, dtype=None):
    return torch.randn(shape, generator=generator, dtype=dtype).to(device)
