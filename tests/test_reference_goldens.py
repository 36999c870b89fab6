"""Parity PIN: tests/golden/ref_model_golden.pt holds outputs of the REFERENCE'S OWN code (its model classes, its LoRA injection, its
trainers' `_compute_loss`, its losses — run through the leaf restatements of tests/shims; generator: tests/golden/make_ref_model_golden.py).

  * CPU: `oracle/mmdit_oracle.py` reproduces every case to fp32 round-off (pred / loss 1e-5, LoRA grads 1e-4) — this is what pins the
    oracle, and with it every "CUDA vs oracle" parity claim, to the reference;
  * CPU: the B200 host path with emulated kernels reproduces the head_dim-128 cases at the bf16 tolerances;
  * GPU (`-m gpu`): the CUDA path reproduces the head_dim-128 cases directly against the reference's vectors (no oracle in between).
Tolerances (relative L2, reference convention tests/src/models/test_qwen_custom.py:550): bf16 B200 path vs fp32 reference: pred 2e-2
(what the reference itself accepts for bf16, test_qwen_custom.py:635-636; measured 4e-3 .. 1.1e-2, the largest with LoRA on every Linear), |loss| 1e-2,
LoRA gradients 2e-2 (3e-2 on the GPU where attention runs in bf16 tensor-core arithmetic)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_common as rc  # noqa: E402

G = torch.load(os.path.join(HERE, "golden", "ref_model_golden.pt"))


def _oracle(spec):
    from oracle import mmdit_oracle as mo
    cfg = spec["cfg"]
    orc = (mo.QwenImageOracle(mo.QwenConfig(**cfg)) if spec["kind"].startswith("qwen") else mo.FluxOracle(mo.FluxConfig(**cfg)))
    mo.add_lora_adapter(orc, r=spec["r"], alpha=spec["alpha"], target_modules=spec["targets"])
    chk = rc.det_fill_(orc, spec["seed"])
    return orc, chk


def _oracle_step(name):
    """(pred, loss) of the oracle's restatement of the trainer recipe on the stored inputs; grads are left on the oracle."""
    from oracle import losses_oracle as lo
    from oracle import mmdit_oracle as mo
    spec, c = rc.CASES[name], G[name]
    orc, chk = _oracle(spec)
    x, ex, kind = c["inputs"], c["extra"], spec["kind"]
    if kind == "qwen":
        crit = None
        if spec.get("loss") == "mask_edit":
            crit = lambda **kw: lo.mask_edit_loss(kw["model_pred"], kw["target"], kw["weighting"], x["edit_mask"], 2.0, 1.0)
        shapes = [[tuple(s) for s in sh] for sh in x["img_shapes"]]
        loss, pred = mo.qwen_compute_loss(orc, x["image_latents"], x["control_latents"], x["prompt_embeds"], x["prompt_embeds_mask"], shapes,
                                          noise=ex["noise"], u=ex["u"], criterion=crit)
    elif kind == "flux":
        loss, pred = mo.flux_compute_loss_shared(orc, x["image_latents"], x["control_latents"], x["pooled_prompt_embeds"], x["prompt_embeds"],
                                                 x["text_ids"], ex["image_ids"], ex["control_ids"], noise=x["noise"], t=x["timestep"])
    else:  # the padded batch exactly as the reference's recipe assembled it (captured model kwargs)
        kw = c["model_kwargs"]
        if kind == "flux_multi":
            out = orc(hidden_states=kw["hidden_states"], encoder_hidden_states=kw["encoder_hidden_states"], pooled_projections=kw["pooled_projections"],
                      timestep=kw["timestep"], img_ids=kw["img_ids"], txt_ids=kw["txt_ids"], guidance=kw["guidance"],
                      attention_mask=kw["attention_mask"])[0]
        else:
            out = orc(hidden_states=kw["hidden_states"], encoder_hidden_states=kw["encoder_hidden_states"], timestep=kw["timestep"],
                      img_shapes=kw["img_shapes"], txt_seq_lens=kw["txt_seq_lens"], attention_mask=kw["attention_mask"],
                      img_offset="reference")[0]
        shapes = x["img_shapes_latent"]
        lt = [s[0][1] * s[0][2] for s in shapes]
        Lt = max(lt)
        noise_in, tmask = torch.zeros(len(lt), Lt, 64), torch.zeros(len(lt), Lt)
        for b, n in enumerate(lt):
            noise_in[b, :n], tmask[b, :n] = x["noise"][b], 1
        pred = out
        loss = lo.attention_mask_mse(out[:, :Lt], noise_in - x["image_latents"][:, :Lt], None, tmask, None)
    return orc, chk, pred, loss


@pytest.mark.parametrize("name", list(rc.CASES))
def test_oracle_matches_reference(name):
    c = G[name]
    orc, chk, pred, loss = _oracle_step(name)
    # same parameter set, same (name-derived) weights as the reference model had
    assert sum(p.numel() for p in orc.parameters()) == c["n_params"]
    assert abs(chk - c["weight_checksum"]) <= 1e-9 * c["weight_checksum"]
    assert sorted(n for n, p in orc.named_parameters() if p.requires_grad) == c["lora_keys"]
    loss.backward()
    ref_pred = c["pred"][:, : pred.shape[1]]  # the trainer recipes slice the target tokens off the model output
    assert pred.shape == ref_pred.shape
    assert rc.rel_l2(pred.detach(), ref_pred) < 1e-5, rc.rel_l2(pred.detach(), ref_pred)
    assert abs(loss.item() - c["loss"].item()) < 1e-5
    for n, p in orc.named_parameters():
        if p.requires_grad:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            ref = c["grads"][n]
            assert (g - ref).norm() <= 1e-4 * ref.norm() + 1e-7, (n, rc.rel_l2(g, ref))


# ------------------------------------------------------------------------------------------------------------------ B200 path
B200_CASES = ["qwen_hd128", "qwen_hd128_plus3", "qwen_hd128_alltargets", "qwen_hd128_editmask", "flux_hd128", "flux_hd128_yaml",
              "flux_custom_multires", "qwen_custom_multires"]


def _b200_model(spec, device, host_only):
    cfg = spec["cfg"]
    if spec["kind"].startswith("qwen"):
        from qflux_b200.qwen_model import QwenB200Config, QwenImageB200
        m = QwenImageB200(QwenB200Config(num_layers=cfg["num_layers"], num_attention_heads=cfg["num_attention_heads"],
                                         joint_attention_dim=cfg["joint_attention_dim"], axes_dims_rope=cfg["axes_dims_rope"]),
                          device=device, _host_only=host_only)
    else:
        from qflux_b200.flux_model import FluxB200, FluxB200Config
        m = FluxB200(FluxB200Config(num_layers=cfg["num_layers"], num_single_layers=cfg["num_single_layers"],
                                    num_attention_heads=cfg["num_attention_heads"], joint_attention_dim=cfg["joint_attention_dim"],
                                    pooled_projection_dim=cfg["pooled_projection_dim"], guidance_embeds=cfg["guidance_embeds"]),
                     device=device, _host_only=host_only)
    m.add_adapter(spec["r"], spec["alpha"], target_modules=spec["targets"])
    # weights by NAME through the public state-dict interface (strict: the key set must be the reference's)
    sd = m.state_dict()
    # det_fill_ works on named_parameters(): wrap the tensors in a throw-away module that exposes them under the same names
    class _Bag(torch.nn.Module):
        def __init__(self, shapes):
            super().__init__()
            self._p = {n: torch.nn.Parameter(torch.empty(s)) for n, s in shapes.items()}

        def named_parameters(self, *a, **k):
            return iter(self._p.items())
    bag = _Bag({n: tuple(t.shape) for n, t in sd.items()})
    chk = rc.det_fill_(bag, spec["seed"])
    m.load_state_dict({n: p.detach() for n, p in bag._p.items()}, strict=True)
    return m, chk


def _b200_step(name, device, host_only, tol_pred, tol_grad):
    from qflux_b200.train_step import FluxKontextStep, QwenImageEditStep
    spec, c = rc.CASES[name], G[name]
    m, chk = _b200_model(spec, device, host_only)
    assert abs(chk - c["weight_checksum"]) <= 1e-9 * c["weight_checksum"], "B200 state-dict key set differs from the reference's"
    x, ex, kind = c["inputs"], c["extra"], spec["kind"]
    bf = lambda t: t.to(device).bfloat16()
    if kind == "qwen":
        step = QwenImageEditStep(m, "mask_edit" if spec.get("loss") == "mask_edit" else "mse")
        emb = dict(image_latents=bf(x["image_latents"]), control_latents=bf(x["control_latents"]), prompt_embeds=bf(x["prompt_embeds"]),
                   img_shapes=[[tuple(s) for s in sh] for sh in x["img_shapes"]])
        if "edit_mask" in x:
            emb["edit_mask"] = x["edit_mask"].to(device)
        loss = step.compute_loss(emb, noise=bf(ex["noise"]), u=ex["u"])
        valid = None
    elif kind == "flux":
        step = FluxKontextStep(m)
        emb = dict(image_latents=bf(x["image_latents"]), control_latents=bf(x["control_latents"]), pooled_prompt_embeds=bf(x["pooled_prompt_embeds"]),
                   prompt_embeds=bf(x["prompt_embeds"]), text_ids=x["text_ids"], image_ids=ex["image_ids"], control_ids=ex["control_ids"])
        loss = step.compute_loss(emb, noise=bf(x["noise"]), t=x["timestep"])
        valid = None
    else:
        shapes = x["img_shapes_latent"]
        lt = [s[0][1] * s[0][2] for s in shapes]
        B, Lt = len(lt), max(lt)
        noise = torch.zeros(B, Lt, 64)
        for b, n in enumerate(lt):
            noise[b, :n] = x["noise"][b]
        if kind == "flux_multi":
            step = FluxKontextStep(m, "attention_mask")
            emb = dict(image_latents=bf(x["image_latents"]), control_latents=bf(x["control_latents"]), pooled_prompt_embeds=bf(x["pooled_prompt_embeds"]),
                       prompt_embeds=bf(x["prompt_embeds"]), text_ids=x["text_ids"], img_shapes=[[tuple(s) for s in sh] for sh in shapes])
            loss = step.compute_loss(emb, noise=bf(noise), t=x["timestep"])
        else:
            step = QwenImageEditStep(m, "attention_mask")
            emb = dict(image_latents=bf(x["image_latents"]), control_latents=bf(x["control_latents"]), prompt_embeds=bf(x["prompt_embeds"]),
                       prompt_embeds_mask=x["prompt_embeds_mask"], img_shapes=[[tuple(s) for s in sh] for sh in shapes])
            # sigma exactly as the stored timestep: u -> idx = (u * 1000).long() -> sigma = (1000 - idx) / 1000
            u = 1.0 - x["timestep"] + 1e-4
            loss = step.compute_loss(emb, noise=bf(noise), u=u)
        valid = lt
    ref = c["pred"]
    pred = m._ws["pred"].view(ref.shape[0], -1, 64).float().cpu().clone()
    loss.backward()
    if valid is None:
        Lq = c["inputs"]["image_latents"].shape[1]
        assert rc.rel_l2(pred[:, :Lq], ref[:, :Lq]) < tol_pred, rc.rel_l2(pred[:, :Lq], ref[:, :Lq])
    else:  # valid target rows of every sample (the reference zeroes padded rows; the training path never reads them)
        for b, n in enumerate(valid):
            assert rc.rel_l2(pred[b, :n], ref[b, :n]) < tol_pred, (b, rc.rel_l2(pred[b, :n], ref[b, :n]))
    assert abs(loss.item() - c["loss"].item()) < 1e-2 * max(1.0, abs(c["loss"].item()))
    num = sum(((p.grad.float().cpu() - c["grads"][n]) ** 2).sum() for n, p in m.named_parameters())
    den = sum((v ** 2).sum() for v in c["grads"].values())
    assert float((num / den).sqrt()) < tol_grad, float((num / den).sqrt())


@pytest.mark.parametrize("name", B200_CASES)
def test_emulated_b200_path_matches_reference(name):
    import emu_lib
    from qflux_b200 import lib
    restore = emu_lib.install(lib)
    try:
        _b200_step(name, "cpu", True, 2e-2, 2e-2)
    finally:
        restore()


@pytest.mark.gpu
@pytest.mark.parametrize("name", B200_CASES)
def test_cuda_path_matches_reference(name):
    _b200_step(name, "cuda", False, 2e-2, 3e-2)
