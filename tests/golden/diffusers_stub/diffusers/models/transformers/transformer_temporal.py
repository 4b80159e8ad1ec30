from dataclasses import dataclass

import torch


@dataclass
class TransformerTemporalModelOutput:
    sample: torch.Tensor
