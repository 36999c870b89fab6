"""qflux_b200 — host side of the B200-native LoRA training step.

Python mirror of the reference's plug points (`trainer.dit` module, `_compute_loss`, loss modules) on top of the
C-ABI library `libqfx_b200.so` (include/qfx.h).  There is no CPU or eager fallback: importing `qflux_b200.lib`
without the built library, or calling any op without a CUDA device, raises.
"""
__version__ = "0.1.0"


def patch_trainer(trainer, use_fused_step: bool = True, **kw):
    """See qflux_b200.integration.patch_trainer (imported lazily: the package itself must import without a GPU)."""
    from .integration import patch_trainer as _p
    return _p(trainer, use_fused_step, **kw)


def from_reference(ref_model, device=None, **kw):
    from .integration import from_reference as _f
    return _f(ref_model, device, **kw)
