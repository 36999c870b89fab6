"""diffusers.models.attention: FeedForward (gelu-approximate) and the processor-plumbing mixins."""
import torch.nn as nn
import torch.nn.functional as F


class GELU(nn.Module):
    """diffusers.models.activations.GELU: proj -> F.gelu(approximate=...)."""

    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states), approximate=self.approximate)


class FeedForward(nn.Module):
    """net = [GELU(dim, 4*dim, tanh), Dropout(0), Linear(4*dim, dim_out)]  (key names net.0.proj / net.2)."""

    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False, inner_dim=None, bias=True):
        super().__init__()
        assert activation_fn == "gelu-approximate", "the shim implements the activation the reference uses"
        inner_dim = int(dim * mult) if inner_dim is None else inner_dim
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.ModuleList([GELU(dim, inner_dim, approximate="tanh", bias=bias), nn.Dropout(dropout),
                                  nn.Linear(inner_dim, dim_out, bias=bias)])

    def forward(self, hidden_states, *args, **kwargs):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class AttentionModuleMixin:
    """diffusers AttentionModuleMixin: processor get / set on an attention module (what the vendored FLUX attention class inherits)."""
    _default_processor_cls = None
    _available_processors = []
    fused_projections = False

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self, return_deprecated_lora=False):
        return self.processor


class AttentionMixin:
    """diffusers AttentionMixin: model-level `attn_processors` / `set_attn_processor` walking the attention sub-modules."""
    @property
    def attn_processors(self):
        return {n + ".processor": m.processor for n, m in self.named_modules() if hasattr(m, "processor")}
