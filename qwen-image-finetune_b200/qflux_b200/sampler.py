"""Validation / inference sampling loop on the fused forward kernels.

Mirrors `sampling_from_embeddings` of the reference trainers — the latent diffusion loop only (encoders and the VAE stay on
the reference side):
  Qwen-Image-Edit  /root/reference/src/qflux/trainer/qwen_image_edit_trainer.py:1116-1289  (true CFG with norm rescaling)
  FLUX-Kontext     /root/reference/src/qflux/trainer/flux_kontext_trainer.py:902-976       (true CFG, plain combination)
and the timestep schedule of `prepare_predict_timesteps` (base_trainer.py:1009-1043): sigmas = linspace(1, 1/N, N), resolution
dependent exponential time shift mu = calculate_shift(image_seq_len) (scheduler/custom_flowmatch_scheduler.py:20-30), Euler
updates x <- x + (sigma_next - sigma) * v of diffusers' FlowMatchEulerDiscreteScheduler (third-party, absent offline: restated
from its published algorithm — set_timesteps with `sigmas`+`mu`, `use_dynamic_shifting=True`, optional `shift_terminal`).
"""
from __future__ import annotations

import math

import torch

BF = torch.bfloat16


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    return image_seq_len * m + (base_shift - m * base_seq_len)


def flow_match_sigmas(num_inference_steps: int, image_seq_len: int, *, base_seq_len=256, max_seq_len=4096, base_shift=0.5,
                      max_shift=1.15, shift_terminal=None) -> torch.Tensor:
    """[N + 1] fp32 sigmas (last = 0); timesteps are sigmas[:-1] * 1000."""
    s = torch.linspace(1.0, 1.0 / num_inference_steps, num_inference_steps, dtype=torch.float64)
    mu = calculate_shift(image_seq_len, base_seq_len, max_seq_len, base_shift, max_shift)
    s = math.exp(mu) / (math.exp(mu) + (1.0 / s - 1.0))  # time_shift_type "exponential", sigma exponent 1.0
    if shift_terminal:
        one_minus = 1.0 - s
        s = 1.0 - one_minus / (one_minus[-1] / (1.0 - shift_terminal))
    return torch.cat([s, torch.zeros(1, dtype=torch.float64)]).float()


def _euler_step(lat, v, sig, i):
    """FlowMatchEulerDiscreteScheduler.step: sample.float() + (sigma_next - sigma) * model_output, cast to the model-output dtype.  The
    step size is a 0-dim fp32 tensor, so its product with the bf16 model output is ROUNDED TO bf16 before the fp32 add (torch type
    promotion) — reproduced here so the trajectories match the reference's scheduler bit for bit."""
    dt = sig[i + 1] - sig[i]
    return (lat.float() + dt * v).to(v.dtype)


def _model_sigma(sigma: float, B: int, device) -> torch.Tensor:
    """What the trainers hand to `dit(timestep=...)`: t = sigma*1000 cast to the weight dtype, then `/ 1000` in that dtype."""
    t = torch.full((B,), sigma * 1000.0, device=device).to(BF)
    return t / 1000


@torch.no_grad()
def sample_qwen(dit, embeddings: dict, *, scheduler_kwargs: dict | None = None) -> torch.Tensor:
    """embeddings: control_latents [B, Lc, 64], prompt_embeds [B, T, J], prompt_embeds_mask [B, T], img_shapes (latent patches),
    latents [B, L, 64] (initial noise), num_inference_steps, true_cfg_scale, optional negative_prompt_embeds(+_mask).
    Returns the final packed latents [B, L, 64] (bf16)."""
    dev = dit.device
    lat = embeddings["latents"].to(dev, BF)
    ctrl = embeddings["control_latents"].to(dev, BF)
    pe, pm = embeddings["prompt_embeds"].to(dev, BF), embeddings["prompt_embeds_mask"].to(dev)
    B, L, _ = lat.shape
    cfg = float(embeddings.get("true_cfg_scale", 1.0))
    do_cfg = cfg > 1 and "negative_prompt_embeds" in embeddings
    if do_cfg:
        ne, nm = embeddings["negative_prompt_embeds"].to(dev, BF), embeddings["negative_prompt_embeds_mask"].to(dev)
    sig = flow_match_sigmas(int(embeddings["num_inference_steps"]), L, **(scheduler_kwargs or {}))
    shapes = embeddings["img_shapes"]
    for i in range(sig.numel() - 1):
        x = torch.cat([lat, ctrl], dim=1)
        ts = _model_sigma(float(sig[i]), B, dev)
        v = dit(hidden_states=x, timestep=ts, encoder_hidden_states=pe, encoder_hidden_states_mask=pm, img_shapes=shapes,
                txt_seq_lens=pm.sum(dim=1).tolist() if pm.device.type == "cpu" else None)[0][:, :L]
        if do_cfg:
            vn = dit(hidden_states=x, timestep=ts, encoder_hidden_states=ne, encoder_hidden_states_mask=nm, img_shapes=shapes,
                     txt_seq_lens=None)[0][:, :L]
            comb = vn + cfg * (v - vn)
            v = comb * (torch.norm(v, dim=-1, keepdim=True) / torch.norm(comb, dim=-1, keepdim=True))
        lat = _euler_step(lat, v, sig, i)
    return lat


@torch.no_grad()
def sample_flux(dit, embeddings: dict, *, scheduler_kwargs: dict | None = None) -> torch.Tensor:
    """embeddings: latents [B, L, 64], latent_ids [L, 3], control_latents, control_ids, pooled_prompt_embeds, prompt_embeds, text_ids,
    guidance (float), num_inference_steps, true_cfg_scale, optional negative_{pooled_prompt_embeds,prompt_embeds,text_ids}."""
    dev = dit.device
    lat = embeddings["latents"].to(dev, BF)
    ctrl = embeddings["control_latents"].to(dev, BF)
    ids = torch.cat([embeddings["latent_ids"].to(dev), embeddings["control_ids"].to(dev)], dim=0)
    pooled, pe, tid = embeddings["pooled_prompt_embeds"].to(dev, BF), embeddings["prompt_embeds"].to(dev, BF), embeddings["text_ids"].to(dev)
    B, L, _ = lat.shape
    cfg = float(embeddings.get("true_cfg_scale", 1.0))
    do_cfg = cfg > 1.0 and "negative_pooled_prompt_embeds" in embeddings
    guidance = torch.full((B,), float(embeddings.get("guidance", 1.0)), device=dev)
    sig = flow_match_sigmas(int(embeddings["num_inference_steps"]), L, **(scheduler_kwargs or {}))
    for i in range(sig.numel() - 1):
        x = torch.cat([lat, ctrl], dim=1)
        ts = _model_sigma(float(sig[i]), B, dev)
        v = dit(hidden_states=x, timestep=ts, guidance=guidance, pooled_projections=pooled, encoder_hidden_states=pe, txt_ids=tid,
                img_ids=ids)[0][:, :L]
        if do_cfg:
            vn = dit(hidden_states=x, timestep=ts, guidance=guidance,
                     pooled_projections=embeddings["negative_pooled_prompt_embeds"].to(dev, BF),
                     encoder_hidden_states=embeddings["negative_prompt_embeds"].to(dev, BF),
                     txt_ids=embeddings["negative_text_ids"].to(dev), img_ids=ids)[0][:, :L]
            v = vn + cfg * (v - vn)
        lat = _euler_step(lat, v, sig, i)
    return lat
