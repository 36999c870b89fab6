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
