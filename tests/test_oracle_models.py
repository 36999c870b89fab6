"""Self-consistency of the MMDiT oracle (parity with diffusers is UNPINNED offline — see oracle header)."""
import torch

from oracle import mmdit_oracle as mo


def tiny_qwen(layers=2):
    cfg = mo.QwenConfig(num_layers=layers, attention_head_dim=64, num_attention_heads=4, joint_attention_dim=512,
                        axes_dims_rope=(8, 28, 28))  # tests/src/models/test_qwen_per_sample_rope.py:243-251
    return mo.init_synthetic_(mo.QwenImageOracle(cfg), std=0.05)


def tiny_flux():
    cfg = mo.FluxConfig(num_layers=2, num_single_layers=1, attention_head_dim=64, num_attention_heads=2,
                        joint_attention_dim=32, pooled_projection_dim=16, axes_dims_rope=(8, 28, 28),
                        guidance_embeds=True)  # tests/src/models/test_flux_per_sample_rope.py:264-278
    return mo.init_synthetic_(mo.FluxOracle(cfg), std=0.05)


def qwen_inputs(B=2, hw=4, T=6, J=512, seed=0):
    g = torch.Generator().manual_seed(seed)
    L = hw * hw
    return dict(image_latents=torch.randn(B, L, 64, generator=g), control_latents=torch.randn(B, L, 64, generator=g),
                prompt_embeds=torch.randn(B, T, J, generator=g), prompt_embeds_mask=torch.ones(B, T, dtype=torch.int64),
                img_shapes=[[(1, hw, hw), (1, hw, hw)]] * B, noise=torch.randn(B, L, 64, generator=g),
                u=torch.rand(B, generator=g))


def test_qwen_tiny_step_and_lora_zero_init():
    m = tiny_qwen()
    x = qwen_inputs()
    base_loss, base_pred = mo.qwen_compute_loss(m, **x)
    mo.add_lora_adapter(m, r=4, alpha=4)  # B = 0  =>  identical output
    loss, pred = mo.qwen_compute_loss(m, **x)
    torch.testing.assert_close(pred, base_pred)
    loss.backward()
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert len(names) == 2 * 4 * 2 and all("lora" in n for n in names)
    gB = [p.grad for n, p in m.named_parameters() if "lora_B" in n]
    gA = [p.grad for n, p in m.named_parameters() if "lora_A" in n]
    assert all(g.abs().sum() > 0 for g in gB) and all(g.abs().sum() == 0 for g in gA)  # dA = 0 when B = 0
    sd = m.state_dict()
    assert "transformer_blocks.0.attn.to_q.lora_A.default.weight" in sd
    assert "transformer_blocks.0.attn.to_q.base_layer.weight" in sd
    assert "transformer_blocks.0.attn.add_q_proj.weight" in sd  # untouched module keeps the diffusers name
    assert "transformer_blocks.1.img_mlp.net.0.proj.weight" in sd and "transformer_blocks.1.txt_mlp.net.2.bias" in sd


def test_qwen_lora_grad_matches_finite_difference():
    m = tiny_qwen(1).double()
    mo.add_lora_adapter(m, r=2, alpha=2, b_std=0.05)
    x = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in qwen_inputs(B=1).items()}
    def f64loss():  # the recipe's loss is fp32 by construction; redo the reduction in fp64 for the FD check
        _, pred = mo.qwen_compute_loss(m, **x)
        return ((pred - (x["noise"] - x["image_latents"])) ** 2).mean()

    f64loss().backward()
    # directional derivative along the gradient itself (the fp32 casts inside RoPE / RMSNorm make single-entry
    # finite differences too noisy): d/de L(theta + e*g) = |g|^2
    ps = [p for p in m.parameters() if p.requires_grad]
    gs = [p.grad.clone() for p in ps]
    g2 = sum((g * g).sum() for g in gs).item()
    eps = 1e-2 / g2 ** 0.5
    with torch.no_grad():
        for p, g in zip(ps, gs):
            p += eps * g
        lp = f64loss()
        for p, g in zip(ps, gs):
            p -= 2 * eps * g
        lm = f64loss()
    fd = (lp - lm).item() / (2 * eps)
    assert abs(fd - g2) <= 2e-3 * g2, (fd, g2)


def test_qwen_rope_layout():
    """QwenEmbedRope (transformer_qwenimage.py:197-254): per-image frame offset, centred h/w, text after max index."""
    r = mo.QwenEmbedRope(10000, [8, 28, 28])
    vid, txt = r([[(1, 4, 6), (1, 4, 6)]], [5, 3], torch.device("cpu"))
    assert vid.shape == (48, 32) and txt.shape == (5, 32) and vid.dtype == torch.complex64
    ang = torch.angle(vid)
    # frame axis (first 4 complex): image 0 -> pos 0 (angle 0), image 1 -> pos 1 (angle = theta^-k)
    assert torch.allclose(ang[:24, :4], torch.zeros(24, 4))
    fr = 1.0 / torch.pow(10000, torch.arange(0, 8, 2).float() / 8)
    assert torch.allclose(ang[24:, :4], fr.expand(24, 4), atol=1e-6)
    # height axis: rows -2,-1,0,1 ; first token of row 0 has h-pos -2 -> angle -2*freq0
    fh = 1.0 / torch.pow(10000, torch.arange(0, 28, 2).float() / 28)
    assert torch.allclose(ang[0, 4:18], -2 * fh, atol=1e-5) and torch.allclose(ang[6 * 2, 4:18], 0 * fh, atol=1e-6)
    # text positions start at max(h//2, w//2) = 3 on all three axes
    assert torch.allclose(torch.angle(txt[0, :4]), torch.remainder(3 * fr + 3.14159265, 2 * 3.14159265) - 3.14159265, atol=1e-4)


def test_flux_tiny_step_config1():
    """BASELINE config 1: tiny 2-layer FLUX (hidden 128) LoRA r=4, 64x64-px latents (16 tokens), one step on CPU."""
    m = tiny_flux()
    mo.add_lora_adapter(m, r=4, alpha=4, b_std=0.05)
    g = torch.Generator().manual_seed(1)
    B, L, T = 1, 16, 8
    ids_img = mo.flux_latent_image_ids(4, 4, 0.0)
    ids_ctl = mo.flux_latent_image_ids(4, 4, 1.0)
    loss, pred = mo.flux_compute_loss_shared(
        m, torch.randn(B, L, 64, generator=g), torch.randn(B, L, 64, generator=g), torch.randn(B, 16, generator=g),
        torch.randn(B, T, 32, generator=g), torch.zeros(T, 3), ids_img, ids_ctl, noise=torch.randn(B, L, 64, generator=g),
        t=torch.tensor([0.5]))
    assert pred.shape == (B, L, 64) and torch.isfinite(loss)
    loss.backward()
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    opt.step()
    grads = [p.grad for n, p in m.named_parameters() if p.requires_grad]
    assert len(grads) == 2 * (4 * 2 + 3 * 1) and all(torch.isfinite(gr).all() and gr.abs().sum() > 0 for gr in grads)


def test_flux_rope_real_equals_complex():
    """FLUX real-valued RoPE == complex multiply on the same angles (ties the two restatements together)."""
    ids = torch.tensor([[0., 1, 2], [1, 3, 0], [0, 0, 5]])
    cos, sin = mo.flux_rope(ids, (8, 28, 28))
    x = torch.randn(1, 3, 2, 64)
    y = mo.apply_rotary_emb_real(x, cos, sin)
    fc = torch.complex(cos[:, 0::2], sin[:, 0::2])
    torch.testing.assert_close(y, mo.apply_rotary_emb_qwen(x, fc), rtol=1e-5, atol=1e-5)


def test_lora_regex_targets():
    m = tiny_flux()
    t = mo.lora_targets(m, r".*single_transformer_blocks\.\d+\.(proj_mlp|proj_out|attn\.to_[qkv])")
    assert sorted(n.split(".", 2)[2] for n, _ in t) == ["attn.to_k", "attn.to_q", "attn.to_v", "proj_mlp", "proj_out"]


# ------------------------------------------------------------------------------------------------------------------
# custom multi-resolution forwards (transformer_qwen_custom.py:384-573, transformer_flux_custom.py:372-741): the self-checks
# the reference's own tests run on them
# ------------------------------------------------------------------------------------------------------------------
def test_qwen_custom_padded_equals_unpadded_per_sample():
    """tests/src/models/test_qwen_per_sample_rope.py:417 (rel-L2 < 1e-4), test_qwen_custom.py:672-692 (padded rows exactly 0)."""
    m = tiny_qwen()
    g = torch.Generator().manual_seed(1)
    shapes = [[(1, 4, 4), (1, 4, 4)], [(1, 2, 4), (1, 4, 2)], [(1, 2, 2), (1, 2, 2)]]
    B, T, J = 3, 6, 512
    L = [sum(f * h * w for f, h, w in sh) for sh in shapes]
    txt = [6, 4, 5]
    Lmax = max(L)
    hs, enc = torch.zeros(B, Lmax, 64), torch.zeros(B, T, J)
    am = torch.zeros(B, T + Lmax, dtype=torch.bool)
    for b in range(B):
        hs[b, :L[b]] = torch.randn(L[b], 64, generator=g)
        enc[b, :txt[b]] = torch.randn(txt[b], J, generator=g)
        am[b, :txt[b]] = True
        am[b, T:T + L[b]] = True
    t = torch.tensor([0.5, 0.25, 0.75])
    with torch.no_grad():
        out = m(hidden_states=hs, encoder_hidden_states=enc, timestep=t, img_shapes=shapes, txt_seq_lens=txt, attention_mask=am)[0]
        out_ref_quirk = m(hidden_states=hs, encoder_hidden_states=enc, timestep=t, img_shapes=shapes, txt_seq_lens=txt, attention_mask=am,
                          img_offset="reference")[0]
        for b in range(B):
            solo = m(hidden_states=hs[b:b + 1, :L[b]], encoder_hidden_states=enc[b:b + 1, :txt[b]], timestep=t[b:b + 1],
                     img_shapes=[shapes[b]], txt_seq_lens=[txt[b]])[0][0]
            assert ((out[b, :L[b]] - solo).norm() / solo.norm()).item() < 1e-4
            assert out[b, L[b]:].abs().max() == 0 if L[b] < Lmax else True
            if txt[b] == T:  # no text padding: the reference's placement of the image table coincides with the aligned one
                torch.testing.assert_close(out_ref_quirk[b], out[b])
            else:            # padded text: the reference applies the image table T - txt_len positions early (documented quirk)
                assert ((out_ref_quirk[b, :L[b]] - solo).norm() / solo.norm()).item() > 1e-3
        # identical shape lists -> the shared path, bit-identical to the stock forward (transformer_qwen_custom.py:458-469)
        same = [shapes[0]] * 2
        a = m(hidden_states=hs[:2], encoder_hidden_states=enc[:2], timestep=t[:2], img_shapes=same, txt_seq_lens=[T, T])[0]
        b_ = m(hidden_states=hs[:2], encoder_hidden_states=enc[:2], timestep=t[:2], img_shapes=[shapes[0]], txt_seq_lens=[T, T])[0]
        assert torch.equal(a, b_)


def test_flux_custom_padded_equals_unpadded_per_sample():
    """tests/src/models/test_flux_per_sample_rope.py:481-486 (rel-L2 < 1e-4), test_flux_transformer_padding.py:91 (atol 1e-5)."""
    m = tiny_flux()
    g = torch.Generator().manual_seed(2)
    hw = [(4, 4), (2, 4), (4, 2)]
    B, T = 3, 5
    L = [2 * h * w for h, w in hw]  # target + one control of the same size
    Lmax = max(L)
    hs, ids = torch.zeros(B, Lmax, 64), torch.zeros(B, Lmax, 3)
    am = torch.ones(B, T + Lmax, dtype=torch.bool)
    for b, (h, w) in enumerate(hw):
        hs[b, :L[b]] = torch.randn(L[b], 64, generator=g)
        ids[b, :L[b]] = torch.cat([mo.flux_latent_image_ids(h, w, 0.0), mo.flux_latent_image_ids(h, w, 1.0)], 0)
        am[b, T + L[b]:] = False
    enc, pooled, t, txt_ids = torch.randn(B, T, 32, generator=g), torch.randn(B, 16, generator=g), torch.tensor([0.5, 0.25, 0.75]), torch.zeros(T, 3)
    with torch.no_grad():
        out = m(hidden_states=hs, encoder_hidden_states=enc, pooled_projections=pooled, timestep=t, img_ids=ids, txt_ids=txt_ids,
                guidance=torch.ones(B), attention_mask=am)[0]
        for b in range(B):
            solo = m(hidden_states=hs[b:b + 1, :L[b]], encoder_hidden_states=enc[b:b + 1], pooled_projections=pooled[b:b + 1],
                     timestep=t[b:b + 1], img_ids=ids[b, :L[b]], txt_ids=txt_ids, guidance=torch.ones(1))[0][0]
            assert ((out[b, :L[b]] - solo).norm() / solo.norm()).item() < 1e-4
            torch.testing.assert_close(out[b, :L[b]], solo, atol=1e-5, rtol=1e-4)
            assert L[b] == Lmax or out[b, L[b]:].abs().max() == 0
