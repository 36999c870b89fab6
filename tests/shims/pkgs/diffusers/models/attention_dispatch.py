"""diffusers.models.attention_dispatch.dispatch_attention_fn, native backend: F.scaled_dot_product_attention on [B,H,S,d]."""
import torch.nn.functional as F


def dispatch_attention_fn(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False,
                          attention_kwargs=None, *, backend=None, parallel_config=None):
    """diffusers.models.attention_dispatch.dispatch_attention_fn, native backend: inputs [B, S, H, d] -> F.scaled_dot_product_attention on
    [B, H, S, d] with the optional additive / boolean mask -> back to [B, S, H, d]."""
    assert backend is None and parallel_config is None, "the shim implements the default (native SDPA) backend only"
    q, k, v = (x.permute(0, 2, 1, 3) for x in (query, key, value))
    out = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale)
    return out.permute(0, 2, 1, 3)
