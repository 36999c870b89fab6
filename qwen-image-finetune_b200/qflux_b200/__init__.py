"""qflux_b200 — host side of the B200-native LoRA training step.

Python mirror of the reference's plug points (`trainer.dit` module, `_compute_loss`, loss modules) on top of the
C-ABI library `libqfx_b200.so` (include/qfx.h).  There is no CPU or eager fallback: importing `qflux_b200.lib`
without the built library, or calling any op without a CUDA device, raises.
"""
__version__ = "0.1.0"
