"""world_size-2 `gloo` test of the data-parallel path: every rank runs the fused step on its own batch, the flat
LoRA-gradient buffer is all-reduced once, and every rank ends with the MEAN gradient (then clip + optimizer step keep the
replicas bit-identical).  The CUDA kernels are emulated (tests/emu_lib.py); what is under test is the host-side exchange step."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "qwen-image-finetune_b200"))
    import emu_lib
    from qflux_b200 import lib
    emu_lib.install(lib)
    from qflux_b200.qwen_model import QwenB200Config, QwenImageB200
    from qflux_b200.train_step import QwenImageEditStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7)  # identical frozen weights + adapters on every rank
    m = QwenImageB200(QwenB200Config(num_layers=1, num_attention_heads=2, joint_attention_dim=128), device="cpu", _host_only=True)
    for k, t in m.w.items():
        if k.endswith("_w") and t.ndim >= 2:
            t.copy_((torch.randn(t.shape) * 0.05).bfloat16())
    m.add_adapter(4, 4, b_std=0.05)
    step = QwenImageEditStep(m, max_grad_norm=0.0)
    g = torch.Generator().manual_seed(100 + rank)  # per-rank data
    rn = lambda *s: torch.randn(*s, generator=g).bfloat16()
    emb = dict(image_latents=rn(1, 16, 64), control_latents=rn(1, 16, 64), prompt_embeds=rn(1, 8, 128), img_shapes=[[(1, 4, 4), (1, 4, 4)]])
    noise, u = rn(1, 16, 64), torch.tensor([0.5])
    # local gradient (no exchange): run the fused step with the process group temporarily hidden
    step._run(*step._prepare(emb, noise, u))
    local = m.G32.clone()
    opt = torch.optim.SGD(list(m.parameters()), lr=0.1)
    step.train_step(emb, optimizer=opt, noise=noise, u=u)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = sum(gathered) / world
    ok_mean = torch.allclose(m.G16.float(), mean.bfloat16().float(), rtol=2e-2, atol=1e-6)
    flat = torch.cat([p.detach().float().reshape(-1) for p in m.parameters()])
    allp = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(allp, flat)
    ok_sync = all(torch.equal(allp[0], t) for t in allp)
    differs = not torch.allclose(gathered[0], gathered[1])
    if rank == 0:
        torch.save(dict(ok_mean=ok_mean, ok_sync=ok_sync, differs=differs), out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp2_gradient_exchange(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["differs"], "ranks must see different batches"
    assert r["ok_mean"], "every rank must hold the mean LoRA gradient after the exchange step"
    assert r["ok_sync"], "replicas must stay bit-identical after the optimizer step"


def _worker_sharded(rank, world, port, out):
    """BASELINE config 4 semantics on 2 ranks: frozen block weights sharded 1/world (all-gathered block by block into a two-slot
    ring), LoRA replicated, gradients all-reduced — must reproduce the un-sharded model's loss and gradients exactly."""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "qwen-image-finetune_b200"))
    import emu_lib
    from qflux_b200 import lib
    emu_lib.install(lib)
    from qflux_b200.qwen_model import QwenB200Config, QwenImageB200
    from qflux_b200.train_step import QwenImageEditStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def build():
        torch.manual_seed(7)
        m = QwenImageB200(QwenB200Config(num_layers=3, num_attention_heads=2, joint_attention_dim=128), device="cpu", _host_only=True)
        for k, t in m.w.items():
            if k.endswith("_w") and t.ndim >= 2:
                t.copy_((torch.randn(t.shape) * 0.05).bfloat16())
            elif k.endswith("_b"):
                t.copy_((torch.randn(t.shape) * 0.02).bfloat16())
        m.add_adapter(4, 4, target_modules=("to_q", "to_out.0", "img_mod.1", "net.2"), b_std=0.05)
        return m

    full, sh = build(), build().shard_frozen_weights()
    per_rank = sh._sharded.shard.numel()
    total = sum(full.w[k].numel() for k in type(full)._PER_LAYER)
    g = torch.Generator().manual_seed(100 + rank)  # per-rank data (data parallel on top of the sharding)
    rn = lambda *s: torch.randn(*s, generator=g).bfloat16()
    emb = dict(image_latents=rn(1, 16, 64), control_latents=rn(1, 16, 64), prompt_embeds=rn(1, 8, 128), img_shapes=[[(1, 4, 4), (1, 4, 4)]])
    noise, u = rn(1, 16, 64), torch.tensor([0.5])
    res = {}
    for name, m in (("full", full), ("sharded", sh)):
        step = QwenImageEditStep(m, max_grad_norm=0.0)
        loss = step._run(*step._prepare(emb, noise, u))
        res[name] = (float(loss), m.G32.clone())
        step._run(*step._prepare(emb, noise, u))  # a second step re-uses the ring (block 0 is resident in slot 0 after the backward)
        assert torch.equal(m.G32, res[name][1])
    # inference path with the ring (two-slot activations, no saved blocks)
    with torch.no_grad():
        kw = dict(hidden_states=rn(1, 32, 64), timestep=torch.tensor([0.5]), encoder_hidden_states=emb["prompt_embeds"],
                  encoder_hidden_states_mask=torch.ones(1, 8, dtype=torch.int64), img_shapes=emb["img_shapes"], txt_seq_lens=[8])
        same_fwd = torch.equal(full(**kw)[0], sh(**kw)[0])
    ok = dict(loss=res["full"][0] == res["sharded"][0], grads=torch.equal(res["full"][1], res["sharded"][1]), fwd=same_fwd,
              shard_fraction=per_rank / total, lora_only_state=all("lora" in k or "transformer_blocks" not in k for k in sh.state_dict()))
    allr = [None] * world
    dist.all_gather_object(allr, ok)
    if rank == 0:
        torch.save(allr, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_frozen_weights_match_replicated(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_sharded, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in torch.load(out):
        assert r["loss"] and r["grads"] and r["fwd"], r
        assert 0.5 <= r["shard_fraction"] < 0.51, "each of 2 ranks keeps half of the block weights (plus alignment padding)"
        assert r["lora_only_state"]


def _worker_sharded_flux(rank, world, port, out):
    """The same sharding for FLUX: one two-slot ring for the double blocks, one for the single blocks; the walk crosses from one
    ring to the other in the forward and back in the backward."""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "qwen-image-finetune_b200"))
    import emu_lib
    from qflux_b200 import lib
    emu_lib.install(lib)
    from qflux_b200.flux_model import FluxB200, FluxB200Config
    from qflux_b200.train_step import FluxKontextStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def build():
        torch.manual_seed(11)
        m = FluxB200(FluxB200Config(num_layers=3, num_single_layers=3, attention_head_dim=128, num_attention_heads=2, joint_attention_dim=64,
                                    pooled_projection_dim=64, guidance_embeds=True), device="cpu", _host_only=True)
        for k, t in m.w.items():
            if k.endswith("_w") and t.ndim >= 2 and "qknorm" not in k:
                t.copy_((torch.randn(t.shape) * 0.05).bfloat16())
            elif k.endswith("_b"):
                t.copy_((torch.randn(t.shape) * 0.02).bfloat16())
        m.add_adapter(4, 4, target_modules=r".*(attn\.to_q|attn\.to_out\.0|norm1\.linear|single_transformer_blocks\.[0-9]+\.(norm\.linear|proj_mlp|proj_out)|ff\.net\.2)",
                      b_std=0.05)
        return m

    full, sh = build(), build().shard_frozen_weights()
    g = torch.Generator().manual_seed(200 + rank)
    rn = lambda *s: torch.randn(*s, generator=g).bfloat16()
    B, hw, T = 1, 4, 8
    L = hw * hw
    emb = dict(image_latents=rn(B, L, 64), control_latents=rn(B, L, 64), pooled_prompt_embeds=rn(B, 64), prompt_embeds=rn(B, T, 64),
               text_ids=torch.zeros(T, 3), image_ids=FluxKontextStep.latent_image_ids(hw, hw, "cpu", 0.0),
               control_ids=FluxKontextStep.latent_image_ids(hw, hw, "cpu", 1.0))
    noise, t = rn(B, L, 64), torch.tensor([0.5])
    res = {}
    for name, m in (("full", full), ("sharded", sh)):
        step = FluxKontextStep(m, max_grad_norm=0.0)
        loss = step._run(*step._prepare(emb, noise, t))
        res[name] = (float(loss), m.G32.clone())
        step._run(*step._prepare(emb, noise, t))  # the second step starts from the rings as the backward left them
        assert torch.equal(m.G32, res[name][1])
    per_rank = sh._sharded.shard.numel() + sh._sharded_s.shard.numel()
    total = sum(full.w[k].numel() for k in type(full)._PER_LAYER + type(full)._PER_SINGLE)
    ok = dict(loss=res["full"][0] == res["sharded"][0], grads=torch.equal(res["full"][1], res["sharded"][1]) and float(res["full"][1].norm()) > 0,
              shard_fraction=per_rank / total, lora_only_state=all("lora" in k or "transformer_blocks" not in k for k in sh.state_dict()))
    allr = [None] * world
    dist.all_gather_object(allr, ok)
    if rank == 0:
        torch.save(allr, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_frozen_weights_flux(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_sharded_flux, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in torch.load(out):
        assert r["loss"] and r["grads"], r
        assert 0.5 <= r["shard_fraction"] < 0.51
        assert r["lora_only_state"]
