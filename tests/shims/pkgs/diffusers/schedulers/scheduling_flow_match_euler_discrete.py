"""diffusers.schedulers.scheduling_flow_match_euler_discrete.FlowMatchEulerDiscreteScheduler — the parts the reference's training
and sampling code touches: the init tables (`timesteps`, `sigmas`), `set_timesteps(sigmas=, mu=)` with exponential dynamic
shifting, `step` (Euler), `scale_noise`.  Karras / exponential / beta sigma variants and stochastic sampling are not restated."""
import math

import numpy as np
import torch

from ..configuration_utils import ConfigMixin, register_to_config


class FlowMatchEulerDiscreteSchedulerOutput:
    """Output record of `step` (prev_sample)."""
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


class FlowMatchEulerDiscreteScheduler(ConfigMixin):
    """diffusers.FlowMatchEulerDiscreteScheduler: init table sigmas = linspace(1, 1/N, N) (static shift applied only without dynamic
    shifting); set_timesteps(sigmas=, mu=) -> exponential time shift e^mu / (e^mu + (1/t - 1)^sigma), sigma = 1, optional stretch to `shift_terminal`,
    timesteps = sigmas * N, a final 0 appended; step: prev = sample.float() + (sigma_next - sigma) * model_output, cast to the
    model_output dtype; scale_noise: sigma * noise + (1 - sigma) * sample."""
    order = 1

    @register_to_config
    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = False,
                 base_shift: float | None = 0.5, max_shift: float | None = 1.15, base_image_seq_len: int | None = 256,
                 max_image_seq_len: int | None = 4096, invert_sigmas: bool = False, shift_terminal: float | None = None,
                 use_karras_sigmas: bool = False, use_exponential_sigmas: bool = False, use_beta_sigmas: bool = False,
                 time_shift_type: str = "exponential", stochastic_sampling: bool = False):
        assert not (use_karras_sigmas or use_exponential_sigmas or use_beta_sigmas or stochastic_sampling or invert_sigmas)
        timesteps = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        timesteps = torch.from_numpy(timesteps).to(dtype=torch.float32)
        sigmas = timesteps / num_train_timesteps
        if not use_dynamic_shifting:  # with dynamic shifting the shift is applied at set_timesteps time, not here
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self._step_index = self._begin_index = None
        self._shift = shift
        self.sigmas = sigmas.to("cpu")
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()

    @classmethod
    def from_config(cls, config, **kwargs):
        return cls(**{**dict(config), **kwargs})

    @property
    def shift(self):
        return self._shift

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index=0):
        self._begin_index = begin_index

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    def time_shift(self, mu, sigma, t):
        if self.config.time_shift_type == "exponential":
            return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)
        return mu / (mu + (1 / t - 1) ** sigma)

    def stretch_shift_to_terminal(self, t):
        one_minus_z = 1 - t
        scale_factor = one_minus_z[-1] / (1 - self.config.shift_terminal)
        return 1 - (one_minus_z / scale_factor)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, timesteps=None):
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError("`mu` must be passed when `use_dynamic_shifting` is set to be `True`")
        assert timesteps is None, "custom timesteps are not restated"
        if sigmas is None:
            ts = np.linspace(self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min), num_inference_steps)
            sigmas = ts / self.config.num_train_timesteps
        else:
            sigmas = np.array(sigmas).astype(np.float32)
            num_inference_steps = len(sigmas)
        self.num_inference_steps = num_inference_steps
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        if self.config.shift_terminal:
            sigmas = self.stretch_shift_to_terminal(sigmas)
        sigmas = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32, device=device)
        timesteps = sigmas * self.config.num_train_timesteps
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self.timesteps = timesteps
        self._step_index = self._begin_index = None

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        if schedule_timesteps is None:
            schedule_timesteps = self.timesteps
        indices = (schedule_timesteps == timestep).nonzero()
        pos = 1 if len(indices) > 1 else 0
        return indices[pos].item()

    def _init_step_index(self, timestep):
        if self.begin_index is None:
            if isinstance(timestep, torch.Tensor):
                timestep = timestep.to(self.timesteps.device)
            self._step_index = self.index_for_timestep(timestep)
        else:
            self._step_index = self._begin_index

    def scale_noise(self, sample, timestep, noise=None):
        sigmas = self.sigmas.to(device=sample.device, dtype=sample.dtype)
        idx = [self.index_for_timestep(t, self.timesteps.to(sample.device)) for t in timestep]
        sigma = sigmas[idx].flatten()
        while len(sigma.shape) < len(sample.shape):
            sigma = sigma.unsqueeze(-1)
        return sigma * noise + (1.0 - sigma) * sample

    def step(self, model_output, timestep, sample, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, generator=None,
             per_token_timesteps=None, return_dict=True):
        """x_{i+1} = x_i + (sigma_{i+1} - sigma_i) * v, computed in fp32 and cast back to the model-output dtype."""
        if self.step_index is None:
            self._init_step_index(timestep)
        sample = sample.to(torch.float32)
        sigma, sigma_next = self.sigmas[self.step_index], self.sigmas[self.step_index + 1]
        prev_sample = sample + (sigma_next - sigma) * model_output
        self._step_index += 1
        prev_sample = prev_sample.to(model_output.dtype)
        if not return_dict:
            return (prev_sample,)
        return FlowMatchEulerDiscreteSchedulerOutput(prev_sample=prev_sample)

    def __len__(self):
        return self.config.num_train_timesteps
