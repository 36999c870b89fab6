"""diffusers.models.normalization: RMSNorm, AdaLayerNormZero(+Single), AdaLayerNormContinuous (layer_norm variants only)."""
import numbers

import torch
import torch.nn as nn


class RMSNorm(nn.Module):
    """var = x.float().pow(2).mean(-1); x = x * rsqrt(var + eps) (fp32 result); if the weight is fp16/bf16 the product is cast to
    the weight dtype BEFORE the multiply; without a weight the result is cast back to the input dtype."""

    def __init__(self, dim, eps: float, elementwise_affine: bool = True, bias: bool = False):
        super().__init__()
        self.eps, self.elementwise_affine = eps, elementwise_affine
        if isinstance(dim, numbers.Integral):
            dim = (dim,)
        self.dim = torch.Size(dim)
        self.weight = self.bias = None
        if elementwise_affine:
            self.weight = nn.Parameter(torch.ones(dim))
            if bias:
                self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, hidden_states):
        input_dtype = hidden_states.dtype
        variance = hidden_states.to(torch.float32).pow(2).mean(-1, keepdim=True)
        hidden_states = hidden_states * torch.rsqrt(variance + self.eps)
        if self.weight is not None:
            if self.weight.dtype in (torch.float16, torch.bfloat16):
                hidden_states = hidden_states.to(self.weight.dtype)
            hidden_states = hidden_states * self.weight
            if self.bias is not None:
                hidden_states = hidden_states + self.bias
        else:
            hidden_states = hidden_states.to(input_dtype)
        return hidden_states


class AdaLayerNormZero(nn.Module):
    """emb = Linear(D, 6D)(SiLU(emb)) -> (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp);
    x = LayerNorm(x; eps 1e-6, no affine) * (1 + scale_msa) + shift_msa."""

    def __init__(self, embedding_dim: int, num_embeddings=None, norm_type="layer_norm", bias=True):
        super().__init__()
        assert num_embeddings is None and norm_type == "layer_norm"
        self.emb = None
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, 6 * embedding_dim, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, timestep=None, class_labels=None, hidden_dtype=None, emb=None):
        emb = self.linear(self.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa, shift_mlp, scale_mlp, gate_mlp


class AdaLayerNormZeroSingle(nn.Module):
    """emb = Linear(D, 3D)(SiLU(emb)) -> (shift_msa, scale_msa, gate_msa); x = LayerNorm(x; eps 1e-6, no affine) * (1 + scale_msa) + shift_msa."""
    def __init__(self, embedding_dim: int, norm_type="layer_norm", bias=True):
        super().__init__()
        assert norm_type == "layer_norm"
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, 3 * embedding_dim, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb=None):
        emb = self.linear(self.silu(emb))
        shift_msa, scale_msa, gate_msa = emb.chunk(3, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa


class AdaLayerNormContinuous(nn.Module):
    """emb = Linear(C, 2D)(SiLU(cond).to(x.dtype)) -> (scale, shift) in THAT order; x = norm(x) * (1 + scale) + shift."""

    def __init__(self, embedding_dim, conditioning_embedding_dim, elementwise_affine=True, eps=1e-5, bias=True, norm_type="layer_norm"):
        super().__init__()
        assert norm_type == "layer_norm"
        self.silu = nn.SiLU()
        self.linear = nn.Linear(conditioning_embedding_dim, embedding_dim * 2, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, eps, elementwise_affine, bias)

    def forward(self, x, conditioning_embedding):
        emb = self.linear(self.silu(conditioning_embedding).to(x.dtype))
        scale, shift = torch.chunk(emb, 2, dim=1)
        return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]
