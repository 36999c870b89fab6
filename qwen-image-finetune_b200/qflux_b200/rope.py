"""Rotary position tables for the joint [text; image] sequence, as (cos, sin) pairs  [S, head_dim/2, 2] fp32.

Semantics follow the reference (product code — does NOT import oracle/):
  Qwen  : /root/reference/src/qflux/models/transformer_qwenimage.py:159-254 (QwenEmbedRope, scale_rope=True):
          image k of a sample sits at frame position k; rows/cols are centred ([-(H-H//2), H//2)); text tokens take
          positions max_vid_index + [0, T) on all three axes, max_vid_index = max_k max(H_k//2, W_k//2).
  FLUX  : /root/reference/src/qflux/models/transformer_flux.py:526-554 (FluxPosEmbed): angle = id * theta^(-2k/dim_axis),
          computed in float64, one (cos, sin) per rotated pair.
Pair i of the table rotates elements (2i, 2i+1) of a head vector — the layout both models use.
"""
from __future__ import annotations

import torch


def _axis_freqs(dim: int, theta: float, dtype) -> torch.Tensor:
    return 1.0 / torch.pow(torch.tensor(float(theta), dtype=dtype), torch.arange(0, dim, 2, dtype=dtype) / dim)


def qwen_rope_table(img_shapes, txt_len: int, axes_dim=(16, 56, 56), theta: float = 10000.0) -> torch.Tensor:
    """img_shapes: [(frame, h, w), ...] of ONE sample (latent-patch units). Returns [txt_len + sum(f*h*w), sum(axes)/2, 2]."""
    fr = [_axis_freqs(d, theta, torch.float32) for d in axes_dim]
    rows = []
    max_vid_index = 0
    for idx, (frame, h, w) in enumerate(img_shapes):
        pf = torch.arange(idx, idx + frame, dtype=torch.float32)
        ph = torch.cat([torch.arange(-(h - h // 2), 0), torch.arange(0, h // 2)]).float()
        pw = torch.cat([torch.arange(-(w - w // 2), 0), torch.arange(0, w // 2)]).float()
        a_f = torch.outer(pf, fr[0]).view(frame, 1, 1, -1).expand(frame, h, w, -1)
        a_h = torch.outer(ph, fr[1]).view(1, h, 1, -1).expand(frame, h, w, -1)
        a_w = torch.outer(pw, fr[2]).view(1, 1, w, -1).expand(frame, h, w, -1)
        rows.append(torch.cat([a_f, a_h, a_w], dim=-1).reshape(frame * h * w, -1))
        max_vid_index = max(h // 2, w // 2, max_vid_index)
    pt = torch.arange(max_vid_index, max_vid_index + txt_len, dtype=torch.float32)
    txt = torch.cat([torch.outer(pt, f) for f in fr], dim=-1)
    ang = torch.cat([txt] + rows, dim=0)
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()


_FREQ_CACHE = {}


def _device_freqs(d: int, theta: float, device) -> torch.Tensor:
    """float64 axis frequencies resident on `device` (uploaded once: the table is rebuilt inside the CUDA-graph-captured step, where a
    host-to-device copy is not allowed)."""
    key = (d, float(theta), str(device))
    if key not in _FREQ_CACHE:
        _FREQ_CACHE[key] = _axis_freqs(d, theta, torch.float64).to(device)
    return _FREQ_CACHE[key]


def flux_rope_table(ids: torch.Tensor, axes_dim=(16, 56, 56), theta: float = 10000.0) -> torch.Tensor:
    """ids: [S, n_axes] (text ids first, as the model concatenates them). Returns [S, sum(axes)/2, 2]."""
    pos = ids.detach().double()  # stays on the ids' device: no host synchronisation inside the training step
    ang = torch.cat([torch.outer(pos[:, i], _device_freqs(d, theta, pos.device)) for i, d in enumerate(axes_dim)], dim=-1)
    return torch.stack([ang.cos().float(), ang.sin().float()], dim=-1).contiguous()
