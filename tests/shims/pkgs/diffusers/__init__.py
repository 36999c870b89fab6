"""Minimal `diffusers` stand-in (tests only) — see tests/shims/README.md.  Only what /root/reference/src/qflux imports."""
__version__ = "0.36.0.shim"


def __getattr__(name):  # heavyweight names resolved lazily so that model-only imports stay light
    if name == "FlowMatchEulerDiscreteScheduler":
        from .schedulers import FlowMatchEulerDiscreteScheduler
        return FlowMatchEulerDiscreteScheduler
    if name in ("FluxKontextPipeline", "QwenImageEditPipeline", "QwenImageEditPlusPipeline", "AutoencoderKL", "AutoencoderKLQwenImage"):
        from .pipelines import LoraSavingPipeline
        return type(name, (LoraSavingPipeline,), {})
    raise AttributeError(name)
