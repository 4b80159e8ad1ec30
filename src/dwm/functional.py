"""Helpers of reference src/dwm/functional.py that the CTSD path uses (:172-193)."""
import torch


def take_sequence_clip(item, start: int, stop: int):
    """Slices the sequence axis (dim 1) of tensors / nested lists; scalars and 1-D
    tensors pass through (reference :172-181)."""
    if isinstance(item, (int, float, bool, str)):
        return item
    if isinstance(item, torch.Tensor):
        return item if len(item.shape) <= 1 else item[:, start:stop]
    if isinstance(item, list):
        assert len(item) > 0 and all(isinstance(i, list) for i in item)
        return [i[start:stop] for i in item]
    raise Exception("Unsupported type to take sequence clip.")


def memory_efficient_split_call(block, tensor: torch.Tensor, func,
                                split_size: int):
    """Applies `func(block, chunk)` over batch chunks (reference :184-193)."""
    if split_size == -1:
        return func(block, tensor)
    return torch.cat([func(block, i) for i in tensor.split(split_size)])
