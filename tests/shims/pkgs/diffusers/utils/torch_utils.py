def maybe_allow_in_graph(cls):
    """diffusers.utils.torch_utils.maybe_allow_in_graph: a torch.compile hint; identity decorator here."""
    return cls


def is_compiled_module(module):
    """diffusers.utils.torch_utils.is_compiled_module: whether the module is a torch.compile OptimizedModule (never, in the tests)."""
    return hasattr(module, "_orig_mod")


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor for a single (or no) generator: torch.randn on the generator's device, then moved."""
    import torch
    rand_device = device
    if generator is not None and hasattr(generator, "device") and generator.device.type != getattr(torch.device(device or "cpu"), "type", "cpu"):
        rand_device = "cpu"
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype, layout=layout or torch.strided).to(device)
