from dataclasses import dataclass

import torch


@dataclass
class Transformer2DModelOutput:
    """diffusers.models.modeling_outputs.Transformer2DModelOutput"""
    sample: torch.Tensor
