"""diffusers.models.attention_processor.Attention — the parameter container + processor dispatch, for the argument set the
reference uses: Attention(query_dim=D, cross_attention_dim=None, added_kv_proj_dim=D, dim_head, heads, out_dim=D,
context_pre_only=False, bias=True, qk_norm="rms_norm", eps) (transformer_qwenimage.py:394-406)."""
import inspect

import torch.nn as nn

from .normalization import RMSNorm


class Attention(nn.Module):
    """diffusers.models.attention_processor.Attention as a PARAMETER CONTAINER: to_q/k/v (+ add_q/k/v_proj, to_add_out, to_out = [Linear, Dropout]
    unless pre_only), norm_q / norm_k (+ norm_added_*) as RMSNorm(dim_head) for qk_norm="rms_norm"; forward delegates to the processor the
    vendored model files install (their processors hold all of the arithmetic)."""
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, kv_heads=None, dim_head=64, dropout=0.0, bias=False,
                 qk_norm=None, added_kv_proj_dim=None, added_proj_bias=True, out_bias=True, eps=1e-5, processor=None, out_dim=None,
                 out_context_dim=None, context_pre_only=None, pre_only=False, elementwise_affine=True):
        super().__init__()
        self.inner_dim = out_dim if out_dim is not None else dim_head * heads
        self.inner_kv_dim = self.inner_dim if kv_heads is None else dim_head * kv_heads
        self.query_dim, self.use_bias = query_dim, bias
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.out_dim = out_dim if out_dim is not None else query_dim
        self.out_context_dim = out_context_dim if out_context_dim is not None else query_dim
        self.context_pre_only, self.pre_only = context_pre_only, pre_only
        self.heads = out_dim // dim_head if out_dim is not None else heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.scale = dim_head ** -0.5
        if qk_norm is None:
            self.norm_q = self.norm_k = None
        elif qk_norm == "rms_norm":
            self.norm_q = RMSNorm(dim_head, eps=eps, elementwise_affine=elementwise_affine)
            self.norm_k = RMSNorm(dim_head, eps=eps, elementwise_affine=elementwise_affine)
        else:
            raise ValueError(f"shim: qk_norm={qk_norm!r} not restated")
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_kv_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_kv_dim, bias=bias)
        if added_kv_proj_dim is not None:
            self.add_k_proj = nn.Linear(added_kv_proj_dim, self.inner_kv_dim, bias=added_proj_bias)
            self.add_v_proj = nn.Linear(added_kv_proj_dim, self.inner_kv_dim, bias=added_proj_bias)
            if context_pre_only is not None:
                self.add_q_proj = nn.Linear(added_kv_proj_dim, self.inner_dim, bias=added_proj_bias)
        if not pre_only:
            self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, self.out_dim, bias=out_bias), nn.Dropout(dropout)])
        if context_pre_only is not None and not context_pre_only:
            self.to_add_out = nn.Linear(self.inner_dim, self.out_context_dim, bias=out_bias)
        if qk_norm is not None and added_kv_proj_dim is not None:
            self.norm_added_q = RMSNorm(dim_head, eps=eps)
            self.norm_added_k = RMSNorm(dim_head, eps=eps)
        else:
            self.norm_added_q = self.norm_added_k = None
        self.processor = processor

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        """kwargs the processor's __call__ does not name are dropped (with a warning in diffusers)."""
        names = set(inspect.signature(self.processor.__call__).parameters.keys())
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in names}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask, **kw)
