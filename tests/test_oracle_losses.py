"""Oracle losses vs golden vectors produced by the reference's own qflux.losses (tests/golden/make_loss_golden.py)."""
import os

import pytest
import torch

from oracle import losses_oracle as lo

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "losses_golden.pt"))


@pytest.mark.parametrize("i", range(len(G["cases"])))
def test_losses_match_reference(i):
    c = G["cases"][i]
    p, t, w, am, em = c["pred"], c["target"], c["weighting"], c["attention_mask"], c["edit_mask"]
    eq = lambda a, b: torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
    eq(lo.mse_loss(p, t, None), c["mse_none_w"])
    eq(lo.mse_loss(p, t, w), c["mse_w"])
    eq(lo.mse_loss(p, t, w, reduction="sum"), c["mse_sum_w"])
    eq(lo.mse_loss(p.bfloat16(), t.bfloat16(), None).float(), c["mse_bf16"])
    eq(lo.mask_edit_loss(p, t, w, None), c["edit_none"])
    eq(lo.mask_edit_loss(p, t, w, em, fg=3.0, bg=0.5), c["edit_w"])
    eq(lo.attention_mask_mse(p, t, w, am, em), c["attn_full"])
    eq(lo.attention_mask_mse(p, t, None, am, None), c["attn_noedit"])
    eq(lo.attention_mask_mse(p, t, None, am, em, reduction="none"), c["attn_none"])


def test_map_mask_to_latent():
    torch.testing.assert_close(lo.map_mask_to_latent(G["img_mask"]), G["latent_mask"])


@pytest.mark.parametrize("i", range(len(G["cases"])))
def test_token_weight_contract(i):
    """The (w, norm) contract of the CUDA flow-loss kernel reproduces every 'mean' loss of the reference."""
    c = G["cases"][i]
    p, t, w, am, em = c["pred"], c["target"], c["weighting"], c["attention_mask"], c["edit_mask"]
    B, T, C = p.shape
    sq = (p - t) ** 2

    def via(kind, **kw):
        tw, norm = lo.token_weights_and_norm(kind, B, T, C, **kw)
        return (sq * tw[..., None]).sum() * norm

    torch.testing.assert_close(via("mse", weighting=w), c["mse_w"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(via("mask_edit", weighting=w, edit_mask=None), c["edit_none"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(via("mask_edit", weighting=w, edit_mask=em, fg=3.0, bg=0.5), c["edit_w"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(via("attention_mask", weighting=w, attention_mask=am, edit_mask=em), c["attn_full"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(via("attention_mask", attention_mask=am), c["attn_noedit"], rtol=1e-5, atol=1e-7)
