"""diffusers.models._modeling_parallel: context-parallel plan records (metadata only; unused on one device)."""
from dataclasses import dataclass


@dataclass
class ContextParallelInput:
    split_dim: int = 0
    expected_dims: int | None = None
    split_output: bool = False


@dataclass
class ContextParallelOutput:
    gather_dim: int = 0
    expected_dims: int | None = None
