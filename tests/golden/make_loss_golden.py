"""Generate tests/golden/losses_golden.pt by importing the REFERENCE's own loss modules.

Run in the build container only (needs /root/reference):
    QFLUX_DOTENV_LOADED=1 PYTHONPATH=/root/reference/src python tests/golden/make_loss_golden.py
The GPU box never runs this; it only reads the committed .pt file.
"""
import os
import sys

import torch

os.environ.setdefault("QFLUX_DOTENV_LOADED", "1")
sys.path.insert(0, "/root/reference/src")
from qflux.losses.attention_mask_loss import AttentionMaskMseLoss  # noqa: E402
from qflux.losses.edit_mask_loss import MaskEditLoss, map_mask_to_latent  # noqa: E402
from qflux.losses.mse_loss import MseLoss  # noqa: E402

g = torch.Generator().manual_seed(20260924)
cases = []
for (B, T, C) in [(1, 16, 64), (2, 24, 8), (4, 32, 64), (3, 7, 5)]:
    pred = torch.randn(B, T, C, generator=g)
    tgt = torch.randn(B, T, C, generator=g)
    wt = torch.rand(B, 1, 1, generator=g) + 0.5
    am = (torch.rand(B, T, generator=g) > 0.3)
    am[:, 0] = True
    em = (torch.rand(B, T, generator=g) > 0.5).float()
    c = dict(pred=pred, target=tgt, weighting=wt, attention_mask=am, edit_mask=em)
    c["mse_none_w"] = MseLoss()(pred, tgt, None)
    c["mse_w"] = MseLoss()(pred, tgt, wt)
    c["mse_sum_w"] = MseLoss(reduction="sum")(pred, tgt, wt)
    c["mse_bf16"] = MseLoss()(pred.bfloat16(), tgt.bfloat16(), None).float()
    c["edit_none"] = MaskEditLoss()(model_pred=pred, target=tgt, weighting=wt, edit_mask=None)
    c["edit_w"] = MaskEditLoss(3.0, 0.5)(model_pred=pred, target=tgt, weighting=wt, edit_mask=em)
    c["attn_full"] = AttentionMaskMseLoss()(model_pred=pred, target=tgt, weighting=wt, attention_mask=am, edit_mask=em)
    c["attn_noedit"] = AttentionMaskMseLoss()(model_pred=pred, target=tgt, weighting=None, attention_mask=am, edit_mask=None)
    c["attn_none"] = AttentionMaskMseLoss(reduction="none")(model_pred=pred, target=tgt, attention_mask=am, edit_mask=em)
    cases.append(c)
img_mask = (torch.rand(2, 64, 96, generator=g) > 0.7).float()
out = dict(cases=cases, img_mask=img_mask, latent_mask=map_mask_to_latent(img_mask))
torch.save(out, os.path.join(os.path.dirname(__file__), "losses_golden.pt"))
print("wrote", len(cases), "cases")
