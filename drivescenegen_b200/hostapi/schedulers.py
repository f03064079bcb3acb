"""``DDPMScheduler`` / ``DDIMScheduler`` with the diffusers 0.20.0 call surface; tensor math in libdsg_b200.

Reference call sites: ``DDPMScheduler()`` DriveSceneGen/scripts/train.py:65; ``.num_train_timesteps`` and ``.add_noise``
DriveSceneGen/pipeline/training_pipeline.py:76,80; ``set_timesteps`` / ``step(...).prev_sample`` through
``DDPMPipeline.__call__`` (training_pipeline.py:26-32, DriveSceneGen/scripts/generation.py:14-20).

Host logic here is the scalar coefficient math, done with fp32 CPU torch scalars in upstream's order of operations
(scheduling_ddpm.py::step, ::_get_variance; scheduling_ddim.py::step — restated in SURVEY.md App. B) and cached as a
per-timestep table on the device.  The elementwise tensor update runs in ``dsg_ddpm_step`` / ``dsg_ddim_step`` /
``dsg_add_noise``.  CUDA tensors only: there is no CPU arithmetic path in the product.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch

from .. import _lib
from .._lib import DsgError, check
from .configuration import ConfigMixin


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """upstream ``utils/torch_utils.randn_tensor`` (CPU generator + CUDA target: draw on CPU, copy)."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        gen_type = generator.device.type
        if gen_type != device.type and gen_type == "cpu":
            rand_device = torch.device("cpu")
        elif gen_type != device.type and gen_type == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gen_type}.")
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _as_f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise DsgError(f"{what}: expected a CUDA tensor; the dsg_b200 schedulers have no CPU arithmetic path")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


class _SchedulerBase(ConfigMixin):
    config_name = "scheduler_config.json"
    order = 1

    def __getattr__(self, name):
        # upstream ConfigMixin.__getattr__: fall back to the config (e.g. scheduler.num_train_timesteps)
        d = self.__dict__.get("_internal_dict")
        if d is not None and name in d:
            return d[name]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def __len__(self):
        return self.config.num_train_timesteps

    def _init_tables(self, num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas):
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._dev_tables: Dict[Tuple, torch.Tensor] = {}
        self._noise_tables: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _leading_timesteps(self, n: int) -> np.ndarray:
        ratio = self.config.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        return ts + self.config.steps_offset if "steps_offset" in self.config else ts

    # -------------------------------------------------------------- add_noise (shared by DDPM and DDIM)
    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor):
        if not original_samples.is_cuda:
            raise DsgError("add_noise: CUDA tensors required (no CPU arithmetic path in the product)")
        dev = original_samples.device
        key = str(dev)
        if key not in self._noise_tables:
            ac = self.alphas_cumprod.to(torch.float32)
            self._noise_tables[key] = ((ac ** 0.5).to(dev), ((1 - ac) ** 0.5).to(dev))
        sa, sb = self._noise_tables[key]
        x0 = _as_f32_cuda(original_samples, "add_noise")
        nz = _as_f32_cuda(noise.to(dev), "add_noise")
        if not timesteps.is_cuda:   # free to check on the host (upstream: IndexError); device tensors are guarded in-kernel
            if timesteps.numel() and (int(timesteps.min()) < 0 or int(timesteps.max()) >= sa.numel()):
                raise IndexError(f"add_noise: timestep outside [0, {sa.numel()})")
        t = timesteps.to(dev, torch.int64).contiguous().flatten()
        batch = x0.shape[0] if t.numel() > 1 else 1
        if t.numel() not in (1, x0.shape[0]):
            raise ValueError("timesteps must have one entry per sample")
        per = x0.numel() // batch
        out = torch.empty_like(x0)
        lib = _lib.load()
        with torch.cuda.device(dev):
            check(lib.dsg_add_noise(x0.data_ptr(), nz.data_ptr(), t.data_ptr(), sa.data_ptr(), sb.data_ptr(),
                                    sa.numel(), out.data_ptr(), batch, per, _stream(dev)), "dsg_add_noise")
        return out


class DDPMScheduler(_SchedulerBase):
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, variance_type: str = "fixed_small",
                 clip_sample: bool = True, prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading", steps_offset: int = 0):
        self.register_to_config(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                beta_schedule=beta_schedule, trained_betas=trained_betas,
                                variance_type=variance_type, clip_sample=clip_sample,
                                prediction_type=prediction_type, thresholding=thresholding,
                                dynamic_thresholding_ratio=dynamic_thresholding_ratio,
                                clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
                                timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        if variance_type != "fixed_small" or prediction_type != "epsilon" or thresholding:
            raise NotImplementedError("dsg_b200 DDPMScheduler supports variance_type='fixed_small', "
                                      "prediction_type='epsilon', thresholding=False (the reference's settings)")
        if timestep_spacing != "leading":
            raise NotImplementedError("only timestep_spacing='leading' (the 0.20.0 default) is supported")
        self._init_tables(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        self.one = torch.tensor(1.0)
        self.custom_timesteps = False
        self.variance_type = variance_type

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, timesteps=None):
        if timesteps is not None:
            raise NotImplementedError("custom timesteps are not supported")
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self.config.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {self.config.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(self._leading_timesteps(num_inference_steps)).to(device)

    def previous_timestep(self, t: int) -> int:
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return t - self.config.num_train_timesteps // n

    def _coef_row(self, t: int):
        """fp32 scalar math in upstream's order (scheduling_ddpm.py::step and ::_get_variance)."""
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        c0 = (a_prev ** 0.5 * cur_beta) / b_t
        cx = cur_alpha ** 0.5 * b_prev / b_t
        var = (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)
        var = torch.clamp(var, min=1e-20)
        sigma = var ** 0.5
        clip = float(self.config.clip_sample_range) if self.config.clip_sample else float("inf")
        return [float(b_t ** 0.5), float(a_t ** 0.5), float(c0), float(cx), float(sigma), clip,
                1.0 if t > 0 else 0.0, 0.0]

    def coef_table(self, device) -> torch.Tensor:
        """device float[num_train_timesteps][8], one row per timestep value, for the current num_inference_steps."""
        key = (str(device), self.num_inference_steps)
        tab = self._dev_tables.get(key)
        if tab is None:
            rows = [self._coef_row(t) for t in range(self.config.num_train_timesteps)]
            tab = torch.tensor(rows, dtype=torch.float32).to(device)
            self._dev_tables[key] = tab
        return tab

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None,
             return_dict: bool = True, variance_noise: Optional[torch.Tensor] = None):
        t = int(timestep)
        if not sample.is_cuda:
            raise DsgError("DDPMScheduler.step: CUDA tensors required (no CPU arithmetic path in the product)")
        dev = sample.device
        eps = _as_f32_cuda(model_output, "step")
        x = _as_f32_cuda(sample, "step")
        z = None
        if t > 0:
            z = variance_noise if variance_noise is not None else randn_tensor(
                model_output.shape, generator=generator, device=dev, dtype=model_output.dtype)
            z = _as_f32_cuda(z, "step")
        prev = torch.empty_like(x)
        lib = _lib.load()
        with torch.cuda.device(dev):
            check(lib.dsg_ddpm_step(eps.data_ptr(), x.data_ptr(), None if z is None else z.data_ptr(),
                                    prev.data_ptr(), x.numel(), self.coef_table(dev).data_ptr(), None, t,
                                    _stream(dev)), "dsg_ddpm_step")
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev)


class DDIMScheduler(_SchedulerBase):
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995,
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False):
        self.register_to_config(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                                set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                prediction_type=prediction_type, thresholding=thresholding,
                                dynamic_thresholding_ratio=dynamic_thresholding_ratio,
                                clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
                                timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr)
        if prediction_type != "epsilon" or thresholding or rescale_betas_zero_snr or timestep_spacing != "leading":
            raise NotImplementedError("dsg_b200 DDIMScheduler supports the 0.20.0 defaults only")
        self._init_tables(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`self.config.train_timesteps`: {self.config.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(self._leading_timesteps(num_inference_steps)).to(device)

    def _coef_row(self, t: int, eta: float):
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        var = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
        std = eta * var ** 0.5
        dir_coef = (1 - a_prev - std ** 2) ** 0.5
        clip = float(self.config.clip_sample_range) if self.config.clip_sample else float("inf")
        return [float(b_t ** 0.5), float(a_t ** 0.5), float(a_prev ** 0.5), float(dir_coef), float(std), clip,
                1.0 if eta > 0 else 0.0, 0.0]

    def coef_table(self, device, eta: float = 0.0) -> torch.Tensor:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        key = (str(device), self.num_inference_steps, float(eta))
        tab = self._dev_tables.get(key)
        if tab is None:
            rows = [self._coef_row(t, eta) for t in range(self.config.num_train_timesteps)]
            tab = torch.tensor(rows, dtype=torch.float32).to(device)
            self._dev_tables[key] = tab
        return tab

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = True):
        if use_clipped_model_output:
            raise NotImplementedError("use_clipped_model_output=True is not supported")
        t = int(timestep)
        if not sample.is_cuda:
            raise DsgError("DDIMScheduler.step: CUDA tensors required (no CPU arithmetic path in the product)")
        dev = sample.device
        eps = _as_f32_cuda(model_output, "step")
        x = _as_f32_cuda(sample, "step")
        z = None
        if eta > 0:
            z = variance_noise if variance_noise is not None else randn_tensor(
                model_output.shape, generator=generator, device=dev, dtype=model_output.dtype)
            z = _as_f32_cuda(z, "step")
        prev = torch.empty_like(x)
        lib = _lib.load()
        with torch.cuda.device(dev):
            check(lib.dsg_ddim_step(eps.data_ptr(), x.data_ptr(), None if z is None else z.data_ptr(),
                                    prev.data_ptr(), x.numel(), self.coef_table(dev, eta).data_ptr(), None, t,
                                    _stream(dev)), "dsg_ddim_step")
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev)
