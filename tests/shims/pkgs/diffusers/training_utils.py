"""diffusers.training_utils: the SD3 timestep-density and loss-weighting helpers, for every scheme they define."""
import math

import torch


def compute_density_for_timestep_sampling(weighting_scheme, batch_size, logit_mean=None, logit_std=None, mode_scale=None,
                                          device="cpu", generator=None):
    """"none" (what the reference passes): u ~ U(0,1) drawn with torch.rand on the CPU from the global RNG."""
    if weighting_scheme == "logit_normal":
        u = torch.normal(mean=logit_mean, std=logit_std, size=(batch_size,), device=device, generator=generator)
        u = torch.nn.functional.sigmoid(u)
    elif weighting_scheme == "mode":
        u = torch.rand(size=(batch_size,), device=device, generator=generator)
        u = 1 - u - mode_scale * (torch.cos(math.pi * u / 2) ** 2 - 1 + u)
    else:
        u = torch.rand(size=(batch_size,), device=device, generator=generator)
    return u


def compute_loss_weighting_for_sd3(weighting_scheme, sigmas=None):
    """diffusers.training_utils.compute_loss_weighting_for_sd3: "sigma_sqrt" -> sigma^-2, "cosmap" -> 2 / (pi * (1 - 2 sigma + 2 sigma^2)),
    anything else (the reference passes "none") -> ones."""
    if weighting_scheme == "sigma_sqrt":
        return (sigmas ** -2.0).float()
    if weighting_scheme == "cosmap":
        bot = 1 - 2 * sigmas + 2 * sigmas ** 2
        return 2 / (math.pi * bot)
    return torch.ones_like(sigmas)
