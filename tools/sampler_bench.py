"""Validation / inference throughput of the fused forward: the reference's `sampling_from_embeddings` loop
(qwen_image_edit_trainer.py:1116-1289 — shifted flow-match sigmas, Euler, true CFG with norm rescale) on `QwenImageB200`, full depth,
512x512 target + 512x512 control, T = 352, 20 steps, true_cfg_scale 4 (two forwards per step).  Beside it: the eager-PyTorch bf16
restatement of the same model (cuBLASLt + SDPA; what the reference would run on this GPU), one forward timed and scaled to the loop.

    python tools/sampler_bench.py [B ...]        (run under gpurun; writes gpurun_out/sampler_bench.json)

The oracle is used here as a BASELINE being measured (like bench.py's library_baseline), never as part of the product path.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qwen-image-finetune_b200"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

BF = torch.bfloat16
STEPS, CFG_SCALE, T, HW = 20, 4.0, 352, 32


def _emb(B, dev, g):
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(BF)
    L = HW * HW
    return dict(latents=rn(B, L, 64), control_latents=rn(B, L, 64), prompt_embeds=rn(B, T, 3584) * 3,
                prompt_embeds_mask=torch.ones(B, T, dtype=torch.int64, device=dev), negative_prompt_embeds=rn(B, T, 3584) * 3,
                negative_prompt_embeds_mask=torch.ones(B, T, dtype=torch.int64, device=dev), img_shapes=[[(1, HW, HW), (1, HW, HW)]] * B,
                num_inference_steps=STEPS, true_cfg_scale=CFG_SCALE)


def ours(B, dev):
    import bench
    from qflux_b200.sampler import sample_qwen
    m = bench.build_model(dev, 60)
    g = torch.Generator(device=dev).manual_seed(3)
    emb = _emb(B, dev, g)
    sample_qwen(m, emb)  # warm-up (workspace, RoPE tables)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = sample_qwen(m, emb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert torch.isfinite(out.float()).all()
    del m
    torch.cuda.empty_cache()
    return dict(B=B, loop_ms=round(ms, 1), forwards=2 * STEPS, ms_per_forward=round(ms / (2 * STEPS), 2), images_per_s=round(B / ms * 1e3, 3))


def eager(B, dev):
    from oracle import mmdit_oracle as mo
    with torch.device(dev):
        m = mo.QwenImageOracle(mo.QwenConfig(num_layers=60))
    m = m.to(BF)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.zero_()
            elif p.ndim == 1:
                p.fill_(1.0)
            else:
                p.normal_(0, 0.02)
    g = torch.Generator(device=dev).manual_seed(3)
    e = _emb(B, dev, g)
    x = torch.cat([e["latents"], e["control_latents"]], 1)
    ts = torch.full((B,), 0.5, device=dev, dtype=BF)

    def fwd():
        with torch.no_grad():
            return m(hidden_states=x, timestep=ts, encoder_hidden_states=e["prompt_embeds"], encoder_hidden_states_mask=e["prompt_embeds_mask"],
                     img_shapes=e["img_shapes"], txt_seq_lens=[T] * B)[0]
    for _ in range(2):
        fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    del m
    torch.cuda.empty_cache()
    return dict(B=B, ms_per_forward=round(ms, 2), loop_ms_scaled=round(ms * 2 * STEPS, 1), images_per_s=round(B / (ms * 2 * STEPS) * 1e3, 3))


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    res = {"config": f"Qwen-Image-Edit predict: 60 blocks, {16 * HW}^2 target + control, T={T}, {STEPS} Euler steps, true CFG {CFG_SCALE} (2 forwards/step), bf16"}
    for B in [int(a) for a in sys.argv[1:]] or [1, 4]:
        res[f"qfx_B{B}"] = ours(B, dev)
        print(json.dumps(res[f"qfx_B{B}"]), flush=True)
        res[f"eager_B{B}"] = eager(B, dev)
        print(json.dumps(res[f"eager_B{B}"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sampler_bench.json"), "w"), indent=1)
