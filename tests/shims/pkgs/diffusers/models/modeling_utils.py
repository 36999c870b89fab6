"""diffusers.models.modeling_utils.ModelMixin: nn.Module + device/dtype properties + gradient-checkpointing switch."""
import functools

import torch
import torch.nn as nn


class ModelMixin(nn.Module):
    """diffusers ModelMixin: an nn.Module with `device` / `dtype` of its first parameter and the gradient-checkpointing switches (the
    loading / saving machinery is not needed: the tests build models from configs)."""
    _supports_gradient_checkpointing = False

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(p for p in self.parameters() if p.is_floating_point()).dtype

    def enable_gradient_checkpointing(self, gradient_checkpointing_func=None):
        """ModelMixin.enable_gradient_checkpointing: torch.utils.checkpoint with use_reentrant=False on every module that has
        a `gradient_checkpointing` attribute."""
        if gradient_checkpointing_func is None:
            gradient_checkpointing_func = functools.partial(torch.utils.checkpoint.checkpoint, use_reentrant=False)
        for m in self.modules():
            if hasattr(m, "gradient_checkpointing"):
                m._gradient_checkpointing_func = gradient_checkpointing_func
                m.gradient_checkpointing = True

    def disable_gradient_checkpointing(self):
        for m in self.modules():
            if hasattr(m, "gradient_checkpointing"):
                m.gradient_checkpointing = False
