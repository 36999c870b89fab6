"""2-GPU check of the sharded-frozen-weights mode (BASELINE config 4 layout) — run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sharded_check.py

Every rank builds the same replicated model and a sharded twin (1/world of each block's frozen weights per rank; a block is assembled in a
two-slot ring on a side stream — once with copy-engine pulls from the peers' IPC-mapped shards, once with an NCCL all-gather), runs the fused training step on per-rank data with both and
compares loss and the flat LoRA gradient (to rounding: the loss sum, LoRA wgrad and dQ use fp32 atomics, so two runs of the SAME
model differ in the last bits too — `self_*` reports that floor); rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen-image-finetune_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def flux_case(dev, rank):
    """FLUX: double-block ring + single-block ring, peer-copy gather, against the replicated model."""
    from qflux_b200.flux_model import FluxB200, FluxB200Config
    from qflux_b200.train_step import FluxKontextStep

    def build():
        g = torch.Generator(device=dev).manual_seed(9)
        m = FluxB200(FluxB200Config(num_layers=3, num_single_layers=4, attention_head_dim=128, num_attention_heads=8, joint_attention_dim=512,
                                    pooled_projection_dim=256, guidance_embeds=True), device=dev)
        for k, t in m.w.items():
            if k.endswith("_w") and t.ndim >= 2 and "qknorm" not in k:
                t.copy_((torch.randn(t.shape, device=dev, generator=g) * 0.03).bfloat16())
            elif k.endswith("_b"):
                t.copy_((torch.randn(t.shape, device=dev, generator=g) * 0.02).bfloat16())
        m.add_adapter(16, 16, target_modules=r".*(attn\.to_[qkv]|attn\.to_out\.0|norm1\.linear|single_transformer_blocks\.[0-9]+\.(norm\.linear|proj_mlp|proj_out)|ff\.net\.2)",
                      b_std=0.05)
        return m

    full, sh = build(), build().shard_frozen_weights(gather="peer")
    g = torch.Generator(device=dev).manual_seed(300 + rank)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    B, hw, T = 2, 16, 40
    L = hw * hw
    emb = dict(image_latents=rn(B, L, 64), control_latents=rn(B, L, 64), pooled_prompt_embeds=rn(B, 256), prompt_embeds=rn(B, T, 512),
               text_ids=torch.zeros(T, 3), image_ids=FluxKontextStep.latent_image_ids(hw, hw, dev, 0.0),
               control_ids=FluxKontextStep.latent_image_ids(hw, hw, dev, 1.0))
    noise, t = rn(B, L, 64), torch.tensor([0.5, 0.25])

    def run(m):
        step = FluxKontextStep(m, max_grad_norm=0.0)
        for _ in range(3):
            loss = step._run(*step._prepare(emb, noise, t))
        torch.cuda.synchronize()
        return float(loss), m.G32.clone()

    rel = lambda a, b: float((a - b).norm() / b.norm())
    (lf, gf), (lf2, gf2), (ls, gs) = run(full), run(full), run(sh)
    r = dict(loss_rel=abs(ls - lf) / abs(lf), grad_rel=rel(gs, gf), self_grad_rel=rel(gf2, gf), grad_norm=float(gf.norm()), gather=sh._sharded.gather)
    r["pass"] = r["loss_rel"] < 1e-5 and r["grad_rel"] < max(1e-4, 10 * r["self_grad_rel"]) and r["gather"] == "peer"
    sh._sharded.close()
    sh._sharded_s.close()
    return r


def main():
    from qflux_b200.qwen_model import QwenB200Config, QwenImageB200
    from qflux_b200.train_step import QwenImageEditStep
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def build():
        g = torch.Generator(device=dev).manual_seed(7)
        m = QwenImageB200(QwenB200Config(num_layers=5, num_attention_heads=8, joint_attention_dim=512), device=dev)
        for k, t in m.w.items():
            if k.endswith("_w") and t.ndim >= 2:
                t.copy_((torch.randn(t.shape, device=dev, generator=g) * 0.03).bfloat16())
            elif k.endswith("_b"):
                t.copy_((torch.randn(t.shape, device=dev, generator=g) * 0.02).bfloat16())
        m.add_adapter(16, 16, target_modules=("to_q", "to_k", "to_v", "to_out.0", "img_mod.1", "net.2"), b_std=0.05)
        return m

    full, sh, sh_nccl = build(), build().shard_frozen_weights(gather="peer"), build().shard_frozen_weights(gather="nccl")
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    B, hw, T = 2, 16, 40
    emb = dict(image_latents=rn(B, hw * hw, 64), control_latents=rn(B, hw * hw, 64), prompt_embeds=rn(B, T, 512) * 3,
               img_shapes=[[(1, hw, hw), (1, hw, hw)]] * B)
    noise, u = rn(B, hw * hw, 64), torch.tensor([0.5, 0.25])
    def run(m):
        step = QwenImageEditStep(m, max_grad_norm=0.0)
        for _ in range(3):  # repeated steps exercise the ring across the forward/backward turn-around
            loss = step._run(*step._prepare(emb, noise, u))
        torch.cuda.synchronize()
        return float(loss), m.G32.clone()

    rel = lambda a, b: float((a - b).norm() / b.norm())
    lf, gf = run(full)
    lf2, gf2 = run(full)
    ls, gs = run(sh)
    ln, gn = run(sh_nccl)
    ok = dict(rank=rank, loss_full=lf, loss_sharded=ls, loss_rel=abs(ls - lf) / abs(lf), grad_rel=rel(gs, gf), self_loss_rel=abs(lf2 - lf) / abs(lf),
              self_grad_rel=rel(gf2, gf), grad_norm=float(gf.norm()), shard_mb=sh._sharded.shard.numel() * 2 / 2 ** 20,
              block_mb=sh._sharded.n_blk * 2 / 2 ** 20)
    ok.update(gather=sh._sharded.gather, nccl_loss_rel=abs(ln - lf) / abs(lf), nccl_grad_rel=rel(gn, gf))
    tol = max(1e-4, 10 * ok["self_grad_rel"])
    ok["pass"] = ok["loss_rel"] < 1e-5 and ok["grad_rel"] < tol and ok["nccl_loss_rel"] < 1e-5 and ok["nccl_grad_rel"] < tol and ok["gather"] == "peer"
    sh._sharded.close()
    ok["flux"] = flux_case(dev, rank)
    ok["pass"] = ok["pass"] and ok["flux"]["pass"]
    allr = [None] * world
    dist.all_gather_object(allr, ok)
    if rank == 0:
        print(json.dumps(dict(ok=all(r["pass"] for r in allr), ranks=allr)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
