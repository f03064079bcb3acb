"""CPU oracle: ``diffusers==0.20.0`` ``DDPMScheduler`` / ``DDIMScheduler`` / ``randn_tensor`` /
``DDPMPipeline.__call__`` restated with plain torch.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Parity PINNED by the upstream known-answer loops
(258.9606 / 0.3372 DDPM, 172.0067 / 0.223967 DDIM; ``tests/test_oracle_kat.py``).

Reference call sites: ``DDPMScheduler()`` DriveSceneGen/scripts/train.py:65; ``.num_train_timesteps``
and ``.add_noise`` DriveSceneGen/pipeline/training_pipeline.py:76,80; ``set_timesteps`` / ``step`` through
``DDPMPipeline.__call__`` DriveSceneGen/pipeline/training_pipeline.py:26-32 and
DriveSceneGen/scripts/generation.py:14-20.  Formulas: /root/repo/SURVEY.md App. B.1-B.3.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """upstream ``utils/torch_utils.randn_tensor``: a CPU generator with a CUDA target draws on the CPU and
    copies; a CUDA generator with a CPU target is an error."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        gen_type = generator.device.type
        if gen_type != device.type and gen_type == "cpu":
            rand_device = torch.device("cpu")
        elif gen_type != device.type and gen_type == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gen_type}.")
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)


class OracleDDPMScheduler:
    """``DDPMScheduler()`` defaults: linear betas 1e-4..0.02 x1000, fixed_small, clip_sample, epsilon, leading."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 clip_sample: bool = True, clip_sample_range: float = 1.0):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def __len__(self):
        return self.num_train_timesteps

    def set_timesteps(self, num_inference_steps: int):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def previous_timestep(self, t):
        n = self.num_inference_steps if self.num_inference_steps else self.num_train_timesteps
        return t - self.num_train_timesteps // n

    def _variance(self, t):
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        cur_beta = 1 - a_t / a_prev
        var = (1 - a_prev) / (1 - a_t) * cur_beta
        return torch.clamp(var, min=1e-20)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None,
             variance_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        t = int(timestep)
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        c0 = (a_prev ** 0.5 * cur_beta) / b_t
        cx = cur_alpha ** 0.5 * b_prev / b_t
        prev = c0 * x0 + cx * sample
        variance = 0
        if t > 0:
            if variance_noise is None:
                variance_noise = randn_tensor(model_output.shape, generator=generator,
                                              device=model_output.device, dtype=model_output.dtype)
            variance = (self._variance(t) ** 0.5) * variance_noise
        return prev + variance

    def add_noise(self, original: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(device=original.device, dtype=original.dtype)
        timesteps = timesteps.to(original.device)
        sa = ac[timesteps] ** 0.5
        sa = sa.flatten()
        while sa.dim() < original.dim():
            sa = sa.unsqueeze(-1)
        sb = (1 - ac[timesteps]) ** 0.5
        sb = sb.flatten()
        while sb.dim() < original.dim():
            sb = sb.unsqueeze(-1)
        return sa * original + sb * noise


class OracleDDIMScheduler:
    """``DDIMScheduler()`` defaults (same beta table, clip_sample, set_alpha_to_one, leading, eta = 0)."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 clip_sample: bool = True, clip_sample_range: float = 1.0, set_alpha_to_one: bool = True):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def __len__(self):
        return self.num_train_timesteps

    def set_timesteps(self, num_inference_steps: int):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             generator=None) -> torch.Tensor:
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        eps = model_output
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        var = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
        std = eta * var ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * eps
        prev = a_prev ** 0.5 * x0 + direction
        if eta > 0:
            z = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                             dtype=model_output.dtype)
            prev = prev + std * z
        return prev

    add_noise = OracleDDPMScheduler.add_noise


@torch.no_grad()
def oracle_ddpm_sample(unet, scheduler, batch_size: int = 1, generator=None, num_inference_steps: int = 1000,
                       sample_size=None, in_channels: int = 3, device="cpu", return_latents: bool = False):
    """``DDPMPipeline.__call__`` (pipelines/ddpm/pipeline_ddpm.py): randn -> loop(unet, step) -> [0,1] NHWC numpy."""
    if isinstance(sample_size, int):
        shape = (batch_size, in_channels, sample_size, sample_size)
    else:
        shape = (batch_size, in_channels, *sample_size)
    image = randn_tensor(shape, generator=generator, device=device)
    scheduler.set_timesteps(num_inference_steps)
    for t in scheduler.timesteps:
        eps = unet(image, t)[0]
        image = scheduler.step(eps, t, image, generator=generator)
    if return_latents:
        return image
    image = (image / 2 + 0.5).clamp(0, 1)
    return image.cpu().permute(0, 2, 3, 1).numpy()
