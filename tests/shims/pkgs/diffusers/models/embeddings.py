"""diffusers.models.embeddings: sinusoidal timestep features, the timestep / guidance / pooled-text conditioning MLPs of FLUX,
and the real-valued rotary helpers FLUX uses."""
import math

import numpy as np
import torch
import torch.nn as nn


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1, scale=1, max_period=10000):
    """freq_i = exp(-ln(max_period) * i / (half - shift)) in fp32; arg = scale * t.float() * freq; [sin | cos], optionally flipped."""
    assert len(timesteps.shape) == 1
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    """diffusers Timesteps: get_timestep_embedding with the stored (num_channels, flip_sin_to_cos, downscale_freq_shift, scale)."""
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float, scale: int = 1):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos = num_channels, flip_sin_to_cos
        self.downscale_freq_shift, self.scale = downscale_freq_shift, scale

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift, scale=self.scale)


class TimestepEmbedding(nn.Module):
    """linear_1 -> SiLU -> linear_2 (act_fn="silu", no conditioning projection)."""

    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu", out_dim=None):
        super().__init__()
        assert act_fn == "silu"
        self.linear_1 = nn.Linear(in_channels, time_embed_dim, True)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim, True)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


class PixArtAlphaTextProjection(nn.Module):
    """diffusers PixArtAlphaTextProjection: linear_1 -> act (gelu_tanh | silu) -> linear_2."""
    def __init__(self, in_features, hidden_size, out_features=None, act_fn="gelu_tanh"):
        super().__init__()
        out_features = hidden_size if out_features is None else out_features
        self.linear_1 = nn.Linear(in_features, hidden_size, True)
        self.act_1 = {"gelu_tanh": nn.GELU(approximate="tanh"), "silu": nn.SiLU()}[act_fn]
        self.linear_2 = nn.Linear(hidden_size, out_features, True)

    def forward(self, caption):
        return self.linear_2(self.act_1(self.linear_1(caption)))


class CombinedTimestepTextProjEmbeddings(nn.Module):
    """temb = MLP_t(sinusoid(t)) + text_embedder(pooled); the sinusoid is cast to the pooled dtype first."""

    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.time_proj = Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0)
        self.timestep_embedder = TimestepEmbedding(in_channels=256, time_embed_dim=embedding_dim)
        self.text_embedder = PixArtAlphaTextProjection(pooled_projection_dim, embedding_dim, act_fn="silu")

    def forward(self, timestep, pooled_projection):
        timesteps_emb = self.timestep_embedder(self.time_proj(timestep).to(dtype=pooled_projection.dtype))
        return timesteps_emb + self.text_embedder(pooled_projection)


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    """temb = (MLP_t(sinusoid(t)) + MLP_g(sinusoid(g))) + text_embedder(pooled)."""

    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.time_proj = Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0)
        self.timestep_embedder = TimestepEmbedding(in_channels=256, time_embed_dim=embedding_dim)
        self.guidance_embedder = TimestepEmbedding(in_channels=256, time_embed_dim=embedding_dim)
        self.text_embedder = PixArtAlphaTextProjection(pooled_projection_dim, embedding_dim, act_fn="silu")

    def forward(self, timestep, guidance, pooled_projection):
        timesteps_emb = self.timestep_embedder(self.time_proj(timestep).to(dtype=pooled_projection.dtype))
        guidance_emb = self.guidance_embedder(self.time_proj(guidance).to(dtype=pooled_projection.dtype))
        return (timesteps_emb + guidance_emb) + self.text_embedder(pooled_projection)


def get_1d_rotary_pos_embed(dim, pos, theta=10000.0, use_real=False, linear_factor=1.0, ntk_factor=1.0,
                            repeat_interleave_real=True, freqs_dtype=torch.float32):
    """freqs = pos (x) theta^(-2k/dim); use_real + repeat_interleave_real: (cos, sin) each repeat_interleave(2) -> fp32 [S, dim]."""
    assert dim % 2 == 0
    if isinstance(pos, int):
        pos = torch.arange(pos)
    if isinstance(pos, np.ndarray):
        pos = torch.from_numpy(pos)
    theta = theta * ntk_factor
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=freqs_dtype, device=pos.device) / dim)) / linear_factor
    freqs = torch.outer(pos, freqs)
    if use_real and repeat_interleave_real:
        return (freqs.cos().repeat_interleave(2, dim=1, output_size=freqs.shape[1] * 2).float(),
                freqs.sin().repeat_interleave(2, dim=1, output_size=freqs.shape[1] * 2).float())
    if use_real:
        return torch.cat([freqs.cos(), freqs.cos()], dim=-1).float(), torch.cat([freqs.sin(), freqs.sin()], dim=-1).float()
    return torch.polar(torch.ones_like(freqs), freqs)


def apply_rotary_emb(x, freqs_cis, use_real=True, use_real_unbind_dim=-1, sequence_dim=2):
    """use_real: pairs (x[2i], x[2i+1]); out = (x.float() * cos + stack(-x_imag, x_real).float() * sin).to(x.dtype)."""
    if use_real:
        cos, sin = freqs_cis
        if sequence_dim == 2:
            cos, sin = cos[None, None, :, :], sin[None, None, :, :]
        elif sequence_dim == 1:
            cos, sin = cos[None, :, None, :], sin[None, :, None, :]
        else:
            raise ValueError(f"`sequence_dim={sequence_dim}` but should be 1 or 2.")
        cos, sin = cos.to(x.device), sin.to(x.device)
        if use_real_unbind_dim == -1:
            x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
            x_rotated = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
        elif use_real_unbind_dim == -2:
            x_real, x_imag = x.reshape(*x.shape[:-1], 2, -1).unbind(-2)
            x_rotated = torch.cat([-x_imag, x_real], dim=-1)
        else:
            raise ValueError(f"`use_real_unbind_dim={use_real_unbind_dim}` but should be -1 or -2.")
        return (x.float() * cos + x_rotated.float() * sin).to(x.dtype)
    x_rotated = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    return torch.view_as_real(x_rotated * freqs_cis.unsqueeze(2)).flatten(3).type_as(x)
