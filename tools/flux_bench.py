"""Throughput of the FLUX-Kontext LoRA step at the BASELINE config-3 shape on one GPU (not the headline metric — bench.py is):
19 double + 38 single blocks, D=3072, T=512 text + 2 x 1024 image tokens, batch 2 per GPU, LoRA r=32 on the target regex of
configs/face_seg_flux_kontext_fp16.yaml (every block Linear, the AdaLN linears, x_embedder), fused clip + AdamW.  Synthetic
weights / embeddings.  Prints one JSON line.

    python tools/flux_bench.py [--steps K] [--warmup W] [--default-targets]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen-image-finetune_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--default-targets", action="store_true", help="to_q,to_k,to_v,to_out.0 instead of the YAML regex")
    a = ap.parse_args()
    from model_check import FLUX_YAML_TARGETS
    from qflux_b200 import lib
    from qflux_b200.flux_model import FluxB200, FluxB200Config
    from qflux_b200.optim import FusedLoraAdamW
    from qflux_b200.train_step import FluxKontextStep
    dev = torch.device("cuda", 0)
    m = FluxB200(FluxB200Config(guidance_embeds=True), device=dev)
    g = torch.Generator(device=dev).manual_seed(1234)
    for k, t in m.w.items():
        if k.endswith("_w") and t.ndim >= 2 and "norm" not in k or k in ("norm_out_w",):
            t.normal_(0.0, 0.02, generator=g)
    targets = ("to_q", "to_k", "to_v", "to_out.0") if a.default_targets else FLUX_YAML_TARGETS
    m.add_adapter(32, 32, target_modules=targets, b_std=0.02)
    n_lora = sum(p.numel() for p in m.parameters())
    B, hw, T = 2, 32, 512
    L = hw * hw
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    emb = dict(image_latents=rn(B, L, 64), control_latents=rn(B, L, 64), pooled_prompt_embeds=rn(B, 768), prompt_embeds=rn(B, T, 4096),
               text_ids=torch.zeros(T, 3, device=dev), image_ids=FluxKontextStep.latent_image_ids(hw, hw, dev, 0.0),
               control_ids=FluxKontextStep.latent_image_ids(hw, hw, dev, 1.0))
    step, opt = FluxKontextStep(m), FusedLoraAdamW(m, lr=1e-4)
    for _ in range(a.warmup):
        step.train_step(emb, opt)
    torch.cuda.synchronize()
    n0 = lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step.train_step(emb, opt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps(dict(workload="FLUX-Kontext LoRA r=32 bf16 512x512, 19+38 blocks, T=512, batch 2, 1 GPU",
                          targets="yaml-regex" if not a.default_targets else "to_q,to_k,to_v,to_out.0", lora_params=n_lora,
                          images_per_s=B / (ms / 1e3), ms_per_step=ms, loss=float(loss), launches_per_step=(lib.LAUNCHES - n0) / a.steps,
                          mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)))


if __name__ == "__main__":
    main()
