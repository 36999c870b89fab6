"""diffusers.models._modeling_parallel: context-parallel plan records (metadata only; unused on one device)."""
from dataclasses import dataclass


@dataclass
class ContextParallelInput:
    """diffusers context-parallel plan record (class-level `_cp_plan` metadata of the vendored models; never executed here)."""
    split_dim: int = 0
    expected_dims: int | None = None
    split_output: bool = False


@dataclass
class ContextParallelOutput:
    """diffusers context-parallel plan record (see ContextParallelInput)."""
    gather_dim: int = 0
    expected_dims: int | None = None
