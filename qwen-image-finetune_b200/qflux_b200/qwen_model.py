"""B200-native Qwen-Image MMDiT with fused-LoRA training step — the drop-in for `trainer.dit`.

Mirrors the interface of the reference's `QwenImageTransformer2DModel`
(/root/reference/src/qflux/models/transformer_qwenimage.py:497-672; called at
/root/reference/src/qflux/trainer/qwen_image_edit_trainer.py:827-836): same forward kwargs, same state-dict key
names (diffusers + PEFT `base_layer` / `lora_A.default` / `lora_B.default`), LoRA A/B exposed as nn.Parameters whose
names contain "lora".  Every FLOP runs in libqfx_b200.so (hand-written sm_100a kernels); this file only sequences the
launches and owns the HBM layout:

  * frozen weights live in layer-stacked fused tensors (q|k|v concatenated per stream, both streams adjacent) so one
    grouped tcgen05 GEMM serves the image and text stream of a block;
  * LoRA factors live zero-padded to 64 rows/cols (`A_pad [64, in]`, `B_pad [out, 64]`) so they enter the K loop of
    the base GEMM as one extra k-block; the nn.Parameters are views of those buffers;
  * tokens are stored stream-major: rows [0, B*T) text, rows [B*T, B*T + B*L) image;
  * activations needed by the backward are kept in HBM (no recompute of GEMMs; 180 GB per GPU has room at B=4):
    block inputs, LayerNorm statistics, pre-norm q|k|v, attention output + logsumexp, mid residual, MLP pre-activation.

No CPU path, no eager-PyTorch fallback: every op raises if the tensors are not on a CUDA device.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import lib
from .rope import qwen_rope_table

BF = torch.bfloat16
PAD = 64  # LoRA rank is padded to one 64-wide k-block


@dataclass
class QwenB200Config:
    patch_size: int = 2
    in_channels: int = 64
    out_channels: int = 16
    num_layers: int = 60
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 3584
    guidance_embeds: bool = False
    axes_dims_rope: tuple = (16, 56, 56)


DEFAULT_TARGETS = ("to_q", "to_k", "to_v", "to_out.0")  # /root/reference/src/qflux/data/config.py:315

# (diffusers sub-module name) -> (group, stream, slot).  group in {"qkv","out","up","down"}; stream 0 = image, 1 = text.
_BLOCK_LINEARS = {
    "attn.to_q": ("qkv", 0, 0), "attn.to_k": ("qkv", 0, 1), "attn.to_v": ("qkv", 0, 2),
    "attn.add_q_proj": ("qkv", 1, 0), "attn.add_k_proj": ("qkv", 1, 1), "attn.add_v_proj": ("qkv", 1, 2),
    "attn.to_out.0": ("out", 0, 0), "attn.to_add_out": ("out", 1, 0),
    "img_mlp.net.0.proj": ("up", 0, 0), "txt_mlp.net.0.proj": ("up", 1, 0),
    "img_mlp.net.2": ("down", 0, 0), "txt_mlp.net.2": ("down", 1, 0),
}


class _LoraSite:
    """LoRA factors of one (block, group, stream): padded buffers shared by the 1 or 3 member modules."""

    def __init__(self, n_slots, d_in, d_out_each, device):
        self.n = n_slots
        self.A_pad = torch.zeros(n_slots * PAD, d_in, device=device, dtype=BF)          # rows g*64.. = A of slot g
        self.B_pad = torch.zeros(n_slots * d_out_each, PAD, device=device, dtype=BF)    # rows g*out.. = B of slot g
        self.members = {}  # slot -> (name, A_param, B_param, gA_off, gB_off)
        self.r = 0


class QwenImageB200(nn.Module):
    def __init__(self, cfg: QwenB200Config, device="cuda", _host_only: bool = False):
        """`_host_only=True` (tests of the naming / LoRA-registry logic) allocates on `device` without requiring CUDA;
        every compute entry point still raises on non-CUDA tensors."""
        super().__init__()
        if not _host_only and not torch.cuda.is_available():
            raise lib.QfxError("QwenImageB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        assert cfg.attention_head_dim == 128, "the sm_100a attention kernels are specialised for head_dim 128"
        self.config = cfg
        self.dev = torch.device(device)
        D = cfg.num_attention_heads * cfg.attention_head_dim
        self.D, self.H, self.L, self.J = D, cfg.num_attention_heads, cfg.num_layers, cfg.joint_attention_dim
        self.C_in, self.C_out = cfg.in_channels, cfg.patch_size ** 2 * cfg.out_channels
        L, J = self.L, self.J
        z = lambda *s: torch.zeros(*s, device=self.dev, dtype=BF)
        # frozen weights (buffers, not Parameters: they never receive gradients on this path)
        self.w = {
            "img_in_w": z(D, self.C_in), "img_in_b": z(D), "txt_norm_w": torch.ones(J, device=self.dev, dtype=BF),
            "txt_in_w": z(D, J), "txt_in_b": z(D),
            "t1_w": z(D, 256), "t1_b": z(D), "t2_w": z(D, D), "t2_b": z(D),
            "mod_w": z(L, 2, 6 * D, D), "mod_b": z(L, 2, 6 * D),
            "qkv_w": z(L, 2, 3 * D, D), "qkv_b": z(L, 2, 3 * D),
            "out_w": z(L, 2, D, D), "out_b": z(L, 2, D),
            "up_w": z(L, 2, 4 * D, D), "up_b": z(L, 2, 4 * D),
            "down_w": z(L, 2, D, 4 * D), "down_b": z(L, 2, D),
            "qknorm_w": torch.ones(L, 4, 128, device=self.dev, dtype=BF),
            "norm_out_w": z(2 * D, D), "norm_out_b": z(2 * D),
            "proj_out_w": z(self.C_out, D), "proj_out_b": z(self.C_out),
        }
        self.sites: dict[tuple[int, str, int], _LoraSite] = {}
        self._lora_params: dict[str, nn.Parameter] = {}
        self.lora_scaling = 0.0
        self.lora_rank = 0
        self.G32 = None  # flat fp32 LoRA-gradient accumulator (what the all-reduce moves)
        self.G16 = None  # flat bf16 gradients (param.grad views)
        self._ws = None
        self._ws_key = None
        self._rope_cache = {}
        self.gradient_checkpointing = False  # accepted for interface compatibility; HBM holds the activations

    # ------------------------------------------------------------------------------------------------ names / state
    @property
    def device(self):
        return self.dev

    @property
    def dtype(self):
        return BF

    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True

    def _weight_views(self) -> dict[str, torch.Tensor]:
        w, D, out = self.w, self.D, {}
        out["img_in.weight"], out["img_in.bias"] = w["img_in_w"], w["img_in_b"]
        out["txt_norm.weight"] = w["txt_norm_w"]
        out["txt_in.weight"], out["txt_in.bias"] = w["txt_in_w"], w["txt_in_b"]
        p = "time_text_embed.timestep_embedder."
        out[p + "linear_1.weight"], out[p + "linear_1.bias"] = w["t1_w"], w["t1_b"]
        out[p + "linear_2.weight"], out[p + "linear_2.bias"] = w["t2_w"], w["t2_b"]
        out["norm_out.linear.weight"], out["norm_out.linear.bias"] = w["norm_out_w"], w["norm_out_b"]
        out["proj_out.weight"], out["proj_out.bias"] = w["proj_out_w"], w["proj_out_b"]
        for l in range(self.L):
            b = f"transformer_blocks.{l}."
            for s, nm in ((0, "img_mod.1"), (1, "txt_mod.1")):
                out[b + nm + ".weight"], out[b + nm + ".bias"] = w["mod_w"][l, s], w["mod_b"][l, s]
            for i, nm in enumerate(("attn.norm_q", "attn.norm_k", "attn.norm_added_q", "attn.norm_added_k")):
                out[b + nm + ".weight"] = w["qknorm_w"][l, i]
            for nm, (grp, s, slot) in _BLOCK_LINEARS.items():
                W, Bv = w[grp + "_w"][l, s], w[grp + "_b"][l, s]
                if grp == "qkv":
                    W, Bv = W[slot * D:(slot + 1) * D], Bv[slot * D:(slot + 1) * D]
                out[b + nm + ".weight"], out[b + nm + ".bias"] = W, Bv
        return out

    def state_dict(self, *args, **kwargs):
        """diffusers / PEFT key names; LoRA'd modules expose `base_layer.*` and `lora_{A,B}.default.weight`."""
        lora_mods = {n.rsplit(".lora_", 1)[0] for n in self._lora_params}
        sd = {}
        for k, v in self._weight_views().items():
            mod, leaf = k.rsplit(".", 1)
            sd[(mod + ".base_layer." + leaf) if mod in lora_mods else k] = v
        for k, p in self._lora_params.items():
            sd[k] = p.detach()
        return sd

    @torch.no_grad()
    def load_state_dict(self, sd, strict=True, assign=False):
        views = self._weight_views()
        missing, unexpected = [], []
        seen = set()
        for k, v in sd.items():
            kk = k.replace(".base_layer.", ".")
            if kk in views:
                views[kk].copy_(v.to(self.dev, BF))
                seen.add(kk)
            elif k in self._lora_params:
                self._lora_params[k].copy_(v.to(self.dev, BF))
                seen.add(k)
            else:
                unexpected.append(k)
        missing = [k for k in list(views) + list(self._lora_params) if k not in seen]
        if strict and (unexpected or [m for m in missing if "lora" not in m]):
            raise KeyError(f"load_state_dict: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
        return missing, unexpected

    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        for k, p in self._lora_params.items():
            yield (prefix + ("." if prefix else "") + k, p)

    def parameters(self, recurse=True):
        for _, p in self.named_parameters():
            yield p

    # ------------------------------------------------------------------------------------------------ LoRA
    def add_adapter(self, r: int, lora_alpha: float, target_modules=DEFAULT_TARGETS, init_lora_weights="gaussian",
                    seed: int = 0, b_std: float = 0.0):
        """PEFT-equivalent of `dit.add_adapter(LoraConfig(...))` (/root/reference/src/qflux/trainer/base_trainer.py:929-941).
        target_modules: list of name suffixes or one full regex (PEFT matching rules)."""
        if not (1 <= r <= PAD and r in (4, 8, 16, 32, 64)):
            raise NotImplementedError(f"LoRA rank {r}: the fused kernels take r in (4, 8, 16, 32, 64)")
        if self._lora_params:
            raise lib.QfxError("an adapter is already attached")
        D = self.D
        self.lora_rank, self.lora_scaling = r, float(lora_alpha) / r
        g = torch.Generator(device=self.dev).manual_seed(seed)
        dims = {"qkv": (D, D, 3), "out": (D, D, 1), "up": (D, 4 * D, 1), "down": (4 * D, D, 1)}
        wanted = []
        for l in range(self.L):
            for nm in _BLOCK_LINEARS:
                full = f"transformer_blocks.{l}.{nm}"
                if isinstance(target_modules, str):
                    ok = re.fullmatch(target_modules, full) is not None
                else:
                    ok = any(full == t or full.endswith("." + t) for t in target_modules)
                if ok:
                    wanted.append((l, nm, full))
        if not wanted:
            raise lib.QfxError(f"no LoRA-capable module matches target_modules={target_modules!r}")
        if isinstance(target_modules, str):  # a regex may also hit modules outside the fused path: fail loudly
            for k in self._weight_views():
                mod = k.rsplit(".", 1)[0]
                if k.endswith(".weight") and re.fullmatch(target_modules, mod) and not any(mod == w[2] for w in wanted):
                    if mod.split(".")[-1] not in ("norm_q", "norm_k", "norm_added_q", "norm_added_k", "txt_norm"):
                        raise NotImplementedError(f"LoRA on `{mod}` is outside the fused hot path (SURVEY.md §8a a12)")
        off = 0
        for l, nm, full in wanted:
            grp, s, slot = _BLOCK_LINEARS[nm]
            if grp == "down":
                raise NotImplementedError("LoRA on FeedForward net.2 (input gelu(u) is not kept in HBM) — next round")
            d_in, d_out, n_slots = dims[grp]
            key = (l, grp, s)
            site = self.sites.get(key)
            if site is None:
                site = self.sites[key] = _LoraSite(n_slots, d_in, d_out, self.dev)
            site.r = r
            A_view = site.A_pad[slot * PAD: slot * PAD + r]            # [r, in]  contiguous
            B_view = site.B_pad[slot * d_out:(slot + 1) * d_out, :r]   # [out, r] strided view
            if init_lora_weights == "gaussian":
                A_view.copy_(torch.randn(r, d_in, device=self.dev, generator=g) / r)
            else:
                A_view.copy_((torch.rand(r, d_in, device=self.dev, generator=g) * 2 - 1) / math.sqrt(d_in))
            if b_std > 0:
                B_view.copy_(torch.randn(d_out, r, device=self.dev, generator=g) * b_std)
            pA, pB = nn.Parameter(A_view), nn.Parameter(B_view)
            gA_off, gB_off = off, off + r * d_in
            off = gB_off + d_out * r
            site.members[slot] = (full, pA, pB, gA_off, gB_off, d_in, d_out)
            self._lora_params[full + ".lora_A.default.weight"] = pA
            self._lora_params[full + ".lora_B.default.weight"] = pB
        self.G32 = torch.zeros(off, device=self.dev, dtype=torch.float32)
        self.G16 = torch.zeros(off, device=self.dev, dtype=BF)
        self._gnorm_sq = torch.zeros(1, device=self.dev, dtype=torch.float32)
        self._gscratch = torch.zeros(4 * D * PAD, device=self.dev, dtype=torch.float32)
        return self

    def bind_param_grads(self):
        """Point every LoRA `param.grad` at its slice of the flat bf16 gradient buffer (no copies)."""
        r = self.lora_rank
        for site in self.sites.values():
            for slot, (full, pA, pB, ga, gb, d_in, d_out) in site.members.items():
                pA.grad = self.G16[ga: ga + r * d_in].view(r, d_in)
                pB.grad = self.G16[gb: gb + d_out * r].view(d_out, r)

    def lora_grad_views(self):
        """{param name: fp32 view into the flat accumulator} — what the parity tests compare against autograd."""
        out = {}
        r = self.lora_rank
        for site in self.sites.values():
            for slot, (full, pA, pB, ga, gb, d_in, d_out) in site.members.items():
                out[full + ".lora_A.default.weight"] = self.G32[ga: ga + r * d_in].view(r, d_in)
                out[full + ".lora_B.default.weight"] = self.G32[gb: gb + d_out * r].view(d_out, r)
        return out

    # ------------------------------------------------------------------------------------------------ workspace
    def _workspace(self, B, T, Limg, train: bool):
        key = (B, T, Limg, train)
        if self._ws_key == key:
            return self._ws
        self._ws = None
        torch.cuda.empty_cache()
        D, H, L = self.D, self.H, self.L
        Mt, Mi = B * T, B * Limg
        M, S = Mt + Mi, T + Limg
        e = lambda *s, dt=BF: torch.empty(*s, device=self.dev, dtype=dt)
        nsave = L if train else 1
        ws = dict(B=B, T=T, Limg=Limg, S=S, Mt=Mt, Mi=Mi, M=M)
        ws["X"] = e(L + 1 if train else 2, M, D)
        ws["xm"], ws["h"] = e(M, D), e(M, 4 * D)
        ws["Q"], ws["K"], ws["V"] = e(B, H, S, 128), e(B, H, S, 128), e(B, H, S, 128)
        ws["qkv"] = e(nsave, M, 3 * D)
        ws["O"] = e(nsave, M, D)
        ws["lse"] = e(nsave, B, H, S, dt=torch.float32)
        ws["xmid"] = e(nsave, M, D)
        ws["u"] = e(nsave, M, 4 * D)
        ws["stats"] = e(nsave, 4, M, dt=torch.float32)  # mean1, rstd1, mean2, rstd2
        ws["sin"], ws["t1"], ws["temb"] = e(B, 256), e(B, D), e(B, D)
        ws["mods"] = e(B, L * 2 * 6 * D)
        ws["fmod"] = e(B, 2 * D)
        ws["fstats"] = e(2, Mi, dt=torch.float32)
        ws["txt_n"] = e(Mt, self.J)
        ws["hn"] = e(Mi, D)
        ws["pred"] = e(Mi, self.C_out)
        # LoRA forward intermediates  T = s * X A^T  (kept for the dB gradient)
        ws["loraT"] = {}
        for (l, grp, s), site in self.sites.items():
            ws["loraT"][(l, grp, s)] = torch.zeros((Mi if s == 0 else Mt), site.n * PAD, device=self.dev, dtype=BF)
        if train:
            ws["dX"] = e(2, M, D)
            ws["dY"] = e(M, D)
            ws["dbig"] = e(M, 4 * D)
            ws["dqkv"] = e(M, 3 * D)
            ws["dxm"] = e(M, D)
            ws["dO"] = e(M, D)
            ws["dOj"], ws["dK"], ws["dV"] = e(B, H, S, 128), e(B, H, S, 128), e(B, H, S, 128)
            ws["dQ"] = e(B, H, S, 128, dt=torch.float32)
            ws["delta"] = e(B, H, S, dt=torch.float32)
            ws["U"] = e(M, 3 * PAD)
            ws["dhn"] = e(Mi, D)
            ws["dpred"] = e(Mi, self.C_out)
            ws["loss"] = torch.zeros(1, device=self.dev, dtype=torch.float32)
        self._ws, self._ws_key = ws, key
        return ws

    def _rope(self, img_shapes, T):
        key = (tuple(tuple(s) for s in img_shapes), T)
        if key not in self._rope_cache:
            self._rope_cache[key] = qwen_rope_table(img_shapes, T, self.config.axes_dims_rope).to(self.dev)
        return self._rope_cache[key]

    # ------------------------------------------------------------------------------------------------ helpers
    def _mod(self, ws, l, s, j):
        """j-th D-wide chunk (shift1, scale1, gate1, shift2, scale2, gate2) of stream s of block l:  [B, D] view."""
        c0 = ((l * 2 + s) * 6 + j) * self.D
        return ws["mods"][:, c0:c0 + self.D]

    def _rows(self, ws, t, s):
        """rows of a stream-major [M, *] tensor: s=1 text, s=0 image."""
        return t[: ws["Mt"]] if s == 1 else t[ws["Mt"]:]

    def _site(self, l, grp, s):
        return self.sites.get((l, grp, s))

    def _lora_T(self, ws, l, grp, s, X):
        """T = scaling * X @ A_pad^T  for the site (None if the site has no adapter)."""
        site = self._site(l, grp, s)
        if site is None:
            return None, None
        Tb = ws["loraT"][(l, grp, s)]
        lib.gemm([lib.gemm_problem(X, site.A_pad, Tb)], site.n * PAD, X.shape[1], alpha=self.lora_scaling)
        return site, Tb

    def _grouped(self, ws, l, grp, src, dst, N, K, epilogue, src_is_pair=False, **epi):
        """One grouped GEMM over (image, text) rows of block l for weight group `grp` with optional fused LoRA."""
        probs = []
        for s in (0, 1):
            A = self._rows(ws, src, s)
            site, Tb = self._lora_T(ws, l, grp, s, A)
            kw = {}
            if site is not None:
                kw = dict(A2=Tb, B2=site.B_pad, kb2=1)
            for k, v in epi.items():
                if v is None:
                    continue
                if k == "gate":
                    kw["gate"] = v[s]
                    kw["rows_per_batch"] = ws["Limg"] if s == 0 else ws["T"]
                else:
                    kw[k] = self._rows(ws, v, s)
            probs.append(lib.gemm_problem(A, self.w[grp + "_w"][l, s], self._rows(ws, dst, s), bias=self.w[grp + "_b"][l, s], **kw))
        lib.gemm(probs, N, K, epilogue=epilogue, lora_group_n=(self.D if grp == "qkv" and any(p.kb2 for p in probs) else 0))

    # ------------------------------------------------------------------------------------------------ forward
    def _forward_impl(self, hidden_states, encoder_hidden_states, timestep, img_shapes, train: bool):
        lib.require_cuda(hidden_states, encoder_hidden_states, timestep)
        B, Limg, _ = hidden_states.shape
        T = encoder_hidden_states.shape[1]
        D, H, L, w = self.D, self.H, self.L, self.w
        ws = self._workspace(B, T, Limg, train)
        Mt = ws["Mt"]
        shapes0 = img_shapes[0] if isinstance(img_shapes[0], (list, tuple)) and isinstance(img_shapes[0][0], (list, tuple)) else img_shapes
        rope = self._rope(shapes0, T)
        assert rope.shape[0] == ws["S"], f"img_shapes {shapes0} do not add up to {Limg} image tokens"
        ws["rope"] = rope
        X0 = ws["X"][0]
        # --- conditioning: sigma (bf16-rounded, transformer_qwenimage.py:624) -> sinusoid -> MLP -> all modulation vectors
        t32 = timestep.to(BF).float().contiguous()
        lib.timestep_sinusoid(t32, 1000.0, ws["sin"])
        lib.gemv_act(ws["sin"], w["t1_w"], w["t1_b"], ws["t1"], act=0)
        lib.gemv_act(ws["t1"], w["t2_w"], w["t2_b"], ws["temb"], act=1)
        lib.gemv_act(ws["temb"], w["mod_w"].view(L * 2 * 6 * D, D), w["mod_b"].view(-1), ws["mods"], act=1)
        lib.gemv_act(ws["temb"], w["norm_out_w"], w["norm_out_b"], ws["fmod"], act=1)
        # --- embedders
        hs = hidden_states.to(BF).reshape(B * Limg, self.C_in)
        lib.gemm([lib.gemm_problem(hs, w["img_in_w"], X0[Mt:], bias=w["img_in_b"])], D, self.C_in)
        lib.rmsnorm_rows(encoder_hidden_states.to(BF).reshape(Mt, self.J), w["txt_norm_w"], ws["txt_n"])
        lib.gemm([lib.gemm_problem(ws["txt_n"], w["txt_in_w"], X0[:Mt], bias=w["txt_in_b"])], D, self.J)
        # --- blocks
        for l in range(L):
            sv = l if train else 0
            Xin, Xout = (ws["X"][l], ws["X"][l + 1]) if train else (ws["X"][l & 1], ws["X"][(l + 1) & 1])
            st, qkv, O, xmid, u = ws["stats"][sv], ws["qkv"][sv], ws["O"][sv], ws["xmid"][sv], ws["u"][sv]
            for s in (0, 1):
                lib.ln_modulate_fwd(self._rows(ws, Xin, s), self._rows(ws, ws["xm"], s), self._mod(ws, l, s, 0), self._mod(ws, l, s, 1),
                                    Limg if s == 0 else T, self._rows(ws, st[0], s), self._rows(ws, st[1], s))
            self._grouped(ws, l, "qkv", ws["xm"], qkv, 3 * D, D, lib.EPI_BIAS)
            for s in (0, 1):
                lib.qk_norm_rope_fwd(self._rows(ws, qkv, s), w["qknorm_w"][l, 2 * s], w["qknorm_w"][l, 2 * s + 1], rope, ws["Q"],
                                     ws["K"], ws["V"], Limg if s == 0 else T, T if s == 0 else 0)
            lib.attn_fwd(ws["Q"], ws["K"], ws["V"], O[:Mt], O[Mt:], T, ws["lse"][sv])
            self._grouped(ws, l, "out", O, xmid, D, D, lib.EPI_RESID_GATE, resid=Xin,
                          gate=(self._mod(ws, l, 0, 2), self._mod(ws, l, 1, 2)))
            for s in (0, 1):
                lib.ln_modulate_fwd(self._rows(ws, xmid, s), self._rows(ws, ws["xm"], s), self._mod(ws, l, s, 3), self._mod(ws, l, s, 4),
                                    Limg if s == 0 else T, self._rows(ws, st[2], s), self._rows(ws, st[3], s))
            self._grouped(ws, l, "up", ws["xm"], ws["h"], 4 * D, D, lib.EPI_GELU, out2=u)
            self._grouped(ws, l, "down", ws["h"], Xout, D, 4 * D, lib.EPI_RESID_GATE, resid=xmid,
                          gate=(self._mod(ws, l, 0, 5), self._mod(ws, l, 1, 5)))
        Xl = ws["X"][L] if train else ws["X"][L & 1]
        ws["Xlast"] = Xl
        # --- output head: AdaLayerNormContinuous (scale, shift) + proj_out, image stream only
        lib.ln_modulate_fwd(Xl[Mt:], ws["hn"], ws["fmod"][:, D:], ws["fmod"][:, :D], Limg, ws["fstats"][0], ws["fstats"][1])
        lib.gemm([lib.gemm_problem(ws["hn"], w["proj_out_w"], ws["pred"], bias=w["proj_out_b"])], self.C_out, D)
        return ws["pred"].view(B, Limg, self.C_out)

    # ------------------------------------------------------------------------------------------------ backward
    def _lora_bwd(self, ws, l, grp, s, dY, Xsaved, n_out_each):
        """LoRA part of a linear's backward for one site. Returns (U, A_pad, kb2) for the fused dgrad, or Nones.
        dY [M_s, n_slots*out]; Xsaved [M_s, in] is the linear's input (for dA)."""
        site = self._site(l, grp, s)
        if site is None:
            return None
        r, Tb = site.r, ws["loraT"][(l, grp, s)]
        U = self._rows(ws, ws["U"], s)[:, : site.n * PAD]
        for g in range(site.n):
            dYg = dY[:, g * n_out_each:(g + 1) * n_out_each]
            Bg = site.B_pad[g * n_out_each:(g + 1) * n_out_each]
            lib.gemm([lib.gemm_problem(dYg, Bg, U[:, g * PAD:(g + 1) * PAD])], PAD, n_out_each, trans_b=True, alpha=self.lora_scaling)
        # weight gradients on the tensor cores: dB_g[out, r] += dY_g^T T_g  (grouped-diagonal), dA_g[r, in] += U_g^T X
        gB, gA = [], []
        for g in range(site.n):
            if g in site.members:
                full, pA, pB, ga, gb, d_in, d_out = site.members[g]
                gB.append(self.G32[gb:])
                gA.append(self.G32[ga:])
            else:  # slot without an adapter (e.g. LoRA on to_q/to_v only): its zero factors produce zeros -> scratch
                gB.append(self._gscratch)
                gA.append(self._gscratch)
        lib.lora_wgrad_tc(dY, Tb, gB, r, 1, r, mode=1 if site.n > 1 else 0, Dg=n_out_each if site.n > 1 else 0)
        lib.lora_wgrad_tc(Xsaved, U, gA, 1, Xsaved.shape[1], r, mode=0)
        return U, site.A_pad, site.n

    def _dgrad_grouped(self, ws, l, grp, dY, dXout, N, K, n_out_each, Xsaved, epilogue=lib.EPI_BIAS, aux=None):
        """dX = dY . W (+ U . A) for both streams of weight group `grp` (W stored [out, in] = [K_red, N])."""
        probs = []
        for s in (0, 1):
            dYs = self._rows(ws, dY, s)
            lb = self._lora_bwd(ws, l, grp, s, dYs, self._rows(ws, Xsaved, s) if Xsaved is not None else None, n_out_each)
            kw = {}
            if lb is not None:
                kw = dict(A2=lb[0], B2=lb[1], kb2=lb[2])
            if aux is not None:
                kw["aux"] = self._rows(ws, aux, s)
            probs.append(lib.gemm_problem(dYs, self.w[grp + "_w"][l, s], self._rows(ws, dXout, s), **kw))
        lib.gemm(probs, N, K, trans_b=True, epilogue=epilogue)

    def _backward_impl(self, dpred):
        """dpred: [B*Limg, C_out] bf16 gradient of the loss wrt the prediction; accumulates LoRA grads into self.G32."""
        ws = self._ws
        assert ws is not None and self._ws_key[3], "backward needs a training-mode forward first"
        B, T, Limg, Mt, Mi, S = ws["B"], ws["T"], ws["Limg"], ws["Mt"], ws["Mi"], ws["S"]
        D, H, L, w = self.D, self.H, self.L, self.w
        rope = ws["rope"]
        # --- head: proj_out dgrad -> final AdaLN backward.  Text rows of the last block output receive no gradient.
        lib.gemm([lib.gemm_problem(dpred, w["proj_out_w"], ws["dhn"])], D, self.C_out, trans_b=True)
        dX = ws["dX"][L & 1]
        dX[:Mt].zero_()
        ws["dY"][:Mt].zero_()
        lib.ln_modulate_bwd(ws["dhn"], ws["Xlast"][Mt:], ws["fstats"][0], ws["fstats"][1], ws["fmod"][:, :D], Limg, dX[Mt:],
                            gate=self._mod(ws, L - 1, 0, 5), dx_gated=ws["dY"][Mt:])
        for l in range(L - 1, -1, -1):
            Xin = ws["X"][l]
            st, qkv, O, xmid, u = ws["stats"][l], ws["qkv"][l], ws["O"][l], ws["xmid"][l], ws["u"][l]
            dXn = ws["dX"][l & 1]  # gradient wrt this block's input (written at the end)
            # ---- MLP branch: dY = dX * gate2 (already produced by the previous LN-backward)
            self._dgrad_grouped(ws, l, "down", ws["dY"], ws["dbig"], 4 * D, D, D, None, epilogue=lib.EPI_DGELU, aux=u)
            if self._site(l, "up", 0) or self._site(l, "up", 1):  # LoRA input = xm2, recomputed from the statistics
                for s in (0, 1):
                    lib.ln_modulate_fwd(self._rows(ws, xmid, s), self._rows(ws, ws["xm"], s), self._mod(ws, l, s, 3),
                                        self._mod(ws, l, s, 4), Limg if s == 0 else T)
            self._dgrad_grouped(ws, l, "up", ws["dbig"], ws["dxm"], D, 4 * D, 4 * D, ws["xm"])
            # ---- norm2 backward: dXmid = dX + LN_bwd ; also emit dXmid * gate1 for the attention out-projection
            for s in (0, 1):
                lib.ln_modulate_bwd(self._rows(ws, ws["dxm"], s), self._rows(ws, xmid, s), self._rows(ws, st[2], s),
                                    self._rows(ws, st[3], s), self._mod(ws, l, s, 4), Limg if s == 0 else T,
                                    self._rows(ws, dX, s), dres=self._rows(ws, dX, s), gate=self._mod(ws, l, s, 2),
                                    dx_gated=self._rows(ws, ws["dY"], s))
            # ---- attention output projection (LoRA input = O)
            self._dgrad_grouped(ws, l, "out", ws["dY"], ws["dO"], D, D, D, O)
            # ---- attention backward
            for s in (0, 1):
                lib.attn_delta(self._rows(ws, O, s), self._rows(ws, ws["dO"], s), ws["delta"], Limg if s == 0 else T,
                               T if s == 0 else 0, ws["dOj"])
                lib.qk_norm_rope_fwd(self._rows(ws, qkv, s), w["qknorm_w"][l, 2 * s], w["qknorm_w"][l, 2 * s + 1], rope, ws["Q"],
                                     ws["K"], ws["V"], Limg if s == 0 else T, T if s == 0 else 0)
            ws["dQ"].zero_()
            lib.attn_bwd(ws["Q"], ws["K"], ws["V"], ws["dOj"], ws["lse"][l], ws["delta"], ws["dQ"], ws["dK"], ws["dV"])
            for s in (0, 1):
                lib.qk_norm_rope_bwd(ws["dQ"], ws["dK"], ws["dV"], self._rows(ws, qkv, s), w["qknorm_w"][l, 2 * s],
                                     w["qknorm_w"][l, 2 * s + 1], rope, self._rows(ws, ws["dqkv"], s), Limg if s == 0 else T,
                                     T if s == 0 else 0)
            # ---- q|k|v projection dgrad (LoRA input = xm1, recomputed from the saved statistics)
            if self._site(l, "qkv", 0) or self._site(l, "qkv", 1):
                for s in (0, 1):
                    lib.ln_modulate_fwd(self._rows(ws, Xin, s), self._rows(ws, ws["xm"], s), self._mod(ws, l, s, 0),
                                        self._mod(ws, l, s, 1), Limg if s == 0 else T)
            self._dgrad_grouped(ws, l, "qkv", ws["dqkv"], ws["dxm"], D, 3 * D, D, ws["xm"])
            # ---- norm1 backward: dXin = dXmid + LN_bwd ; emit dXin * gate2 of the previous block
            for s in (0, 1):
                has_prev = l > 0
                lib.ln_modulate_bwd(self._rows(ws, ws["dxm"], s), self._rows(ws, Xin, s), self._rows(ws, st[0], s),
                                    self._rows(ws, st[1], s), self._mod(ws, l, s, 1), Limg if s == 0 else T,
                                    self._rows(ws, dXn, s), dres=self._rows(ws, dX, s),
                                    gate=self._mod(ws, l - 1, s, 5) if has_prev else None,
                                    dx_gated=self._rows(ws, ws["dY"], s) if has_prev else None)
            dX = dXn

    # ------------------------------------------------------------------------------------------------ public API
    def forward(self, hidden_states, encoder_hidden_states=None, encoder_hidden_states_mask=None, timestep=None,
                img_shapes=None, txt_seq_lens=None, guidance=None, attention_kwargs=None, return_dict=False):
        """Same signature as the reference model.  With grad enabled and an adapter attached the result carries
        autograd history to the LoRA parameters (custom Function around the fused backward)."""
        if torch.is_grad_enabled() and self._lora_params:
            plist = list(self._lora_params.values())
            out = _ModelFn.apply(self, hidden_states, encoder_hidden_states, timestep, img_shapes, *plist)
        else:
            out = self._forward_impl(hidden_states, encoder_hidden_states, timestep, img_shapes, train=False).clone()
        return (out,)

    def zero_lora_grads(self):
        self.G32.zero_()

    def finalize_grads(self, world_size: int = 1, max_norm: float = 0.0):
        """fp32 accumulator (already all-reduced by the caller) -> mean over ranks -> clip -> bf16 `param.grad`."""
        lib.grad_finalize(self.G32, 1.0 / world_size, max_norm, self._gnorm_sq, self.G16)
        self.bind_param_grads()
        return self._gnorm_sq


class _ModelFn(torch.autograd.Function):
    """Autograd bridge: forward = fused kernels (activations kept in the model workspace), backward = fused kernels
    writing LoRA gradients into the flat accumulator; the returned per-parameter grads are views of it."""

    @staticmethod
    def forward(ctx, model, hidden_states, encoder_hidden_states, timestep, img_shapes, *params):
        ctx.model = model
        pred = model._forward_impl(hidden_states, encoder_hidden_states, timestep, img_shapes, train=True)
        return pred.clone()

    @staticmethod
    def backward(ctx, dpred):
        m = ctx.model
        m.G32.zero_()
        m._backward_impl(dpred.to(BF).reshape(-1, m.C_out).contiguous())
        views = m.lora_grad_views()
        grads = tuple(views[k].to(BF) for k in m._lora_params)
        return (None, None, None, None, None) + grads
