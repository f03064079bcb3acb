"""``DDPMPipeline`` / ``DDIMPipeline`` with the diffusers 0.20.0 call surface.

Reference call sites: ``DDPMPipeline(unet=..., scheduler=...)`` and ``pipeline.save_pretrained(dir)``
DriveSceneGen/pipeline/training_pipeline.py:101,107; ``pipeline(num_inference_steps=750, batch_size=...,
generator=torch.manual_seed(seed), output_type="np.array", return_dict=False)`` training_pipeline.py:26-32;
``DDPMPipeline.from_pretrained(path, variant="fp16").to('cuda')`` and ``ddpm(batch_size=5,
num_inference_steps=750).images`` DriveSceneGen/scripts/generation.py:7,14-20.

The sampling loop keeps upstream's structure (randn -> for t: unet -> scheduler.step -> post-process).  On CUDA the
loop body (U-Net forward + scheduler step) is captured once into a CUDA graph and replayed per step with the timestep
row and the variance noise as device-side inputs; the noise itself is still drawn by ``torch.randn`` with the
caller's generator (RNG parity with upstream ``randn_tensor``).  The post-process ``(x/2+.5).clamp(0,1)`` -> NHWC
(-> uint8 for PIL) runs in ``dsg_latent_to_image``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _lib
from .._lib import DsgError, check
from .configuration import DIFFUSERS_VERSION
from .schedulers import DDIMScheduler, DDPMScheduler, randn_tensor
from .unet2d import UNet2DModel


@dataclass
class ImagePipelineOutput:
    images: Union[List["PIL.Image.Image"], np.ndarray]  # noqa: F821


def numpy_to_pil(images: np.ndarray):
    from PIL import Image
    if images.ndim == 3:
        images = images[None, ...]
    if images.dtype != np.uint8:
        images = (images * 255).round().astype("uint8")
    if images.shape[-1] == 1:
        return [Image.fromarray(im.squeeze(), mode="L") for im in images]
    return [Image.fromarray(im) for im in images]


_SCHEDULERS = {"DDPMScheduler": DDPMScheduler, "DDIMScheduler": DDIMScheduler}


class DenoiseSession:
    """One captured CUDA graph = U-Net forward + scheduler step on static device buffers.

    ``x`` holds the current sample; ``step(t)`` advances it by one denoising step.  ``step_from_host`` is the
    end-to-end form: the step's variance noise comes from (pinned) host memory and the new sample is copied back
    to (pinned) host memory, both inside the call.
    """

    def __init__(self, unet: UNet2DModel, scheduler, shape, ddim: bool = False, eta: float = 0.0):
        self.lib = _lib.load()
        dev = unet.device
        self.dev = dev
        self.x = torch.zeros(shape, dtype=torch.float32, device=dev)       # current sample
        self.x_next = torch.zeros_like(self.x)
        self.eps = torch.zeros_like(self.x)
        self.z = torch.zeros_like(self.x)                                  # variance noise of this step
        self.t_f = torch.zeros(shape[0], dtype=torch.float32, device=dev)  # timestep value (float) per sample
        self.row = torch.zeros(1, dtype=torch.int32, device=dev)           # coefficient-table row (= timestep)
        self.table = scheduler.coef_table(dev, eta) if ddim else scheduler.coef_table(dev)
        eng = unet.engine()
        prog = eng.program(shape[0], shape[2], shape[3])
        step_fn = self.lib.dsg_ddim_step if ddim else self.lib.dsg_ddpm_step

        def body():
            prog.run(self.x, self.t_f, self.eps)
            st = torch.cuda.current_stream(dev).cuda_stream
            check(step_fn(self.eps.data_ptr(), self.x.data_ptr(), self.z.data_ptr(), self.x_next.data_ptr(),
                          self.x.numel(), self.table.data_ptr(), self.row.data_ptr(), 0, st), "scheduler step")
            self.x.copy_(self.x_next)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()  # warm-up outside capture (function attributes, lazy module loading)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()
        self.launches_per_step = prog.n_launches + 2

    def set_step(self, t: int):
        self.t_f.fill_(float(t))
        self.row.fill_(int(t))

    def step(self, t: int, noise: Optional[torch.Tensor] = None):
        self.set_step(t)
        if noise is not None:
            self.z.copy_(noise, non_blocking=True)
        self.graph.replay()
        return self.x

    def step_from_host(self, t: int, noise_host: torch.Tensor, out_host: torch.Tensor):
        self.step(t, noise_host)
        out_host.copy_(self.x, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return out_host

    def _pipe_state(self):
        if getattr(self, "_pipe", None) is None:
            ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]
            self._pipe = {"h2d": torch.cuda.Stream(device=self.dev), "d2h": torch.cuda.Stream(device=self.dev),
                          "z": [torch.zeros_like(self.z), torch.zeros_like(self.z)],
                          "x": [torch.zeros_like(self.x), torch.zeros_like(self.x)],
                          "z_in": ev(), "z_free": ev(), "x_ready": ev(), "x_out": ev()}
        return self._pipe

    def run_from_host(self, timesteps, noise_host, out_host, on_result=None):
        """End-to-end loop with the copies taken off the critical path.

        Step ``i`` takes its variance noise from the pinned host tensor ``noise_host[i % len(noise_host)]`` and lands its
        new sample in the pinned host tensor ``out_host[i % len(out_host)]`` (``len(out_host) >= 2``).  The host->device
        copy of step ``i+1`` and the device->host copy of step ``i-1`` run on their own streams while the graph of step
        ``i`` computes (a step depends on the previous step's sample only on the device); the host waits for result
        ``i-1`` after it has queued step ``i`` and then calls ``on_result(i-1, tensor)``.  Every step's input still
        crosses PCIe and every step's result is still read on the host before the call returns.
        """
        if len(out_host) < 2:
            raise ValueError("run_from_host needs at least two host output buffers")
        ps = self._pipe_state()
        main = torch.cuda.current_stream(self.dev)
        n = len(timesteps)
        nz = len(noise_host)
        started = [False, False]      # has stage slot s been used in this call (events valid)?

        def h2d(i):
            s = i % 2
            with torch.cuda.stream(ps["h2d"]):
                if started[s]:
                    ps["h2d"].wait_event(ps["z_free"][s])
                else:
                    ps["h2d"].wait_stream(main)
                ps["z"][s].copy_(noise_host[i % nz], non_blocking=True)
                ps["z_in"][s].record(ps["h2d"])

        outs = [False, False]
        if n:
            h2d(0)
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                h2d(i + 1)
            main.wait_event(ps["z_in"][s])
            self.z.copy_(ps["z"][s])
            ps["z_free"][s].record(main)
            started[s] = True
            self.set_step(int(timesteps[i]))
            self.graph.replay()
            if outs[s]:
                main.wait_event(ps["x_out"][s])      # the copy out of this stage slot two steps ago has finished
            ps["x"][s].copy_(self.x)
            ps["x_ready"][s].record(main)
            with torch.cuda.stream(ps["d2h"]):
                ps["d2h"].wait_event(ps["x_ready"][s])
                out_host[i % len(out_host)].copy_(ps["x"][s], non_blocking=True)
                ps["x_out"][s].record(ps["d2h"])
            outs[s] = True
            if i >= 1:
                ps["x_out"][(i - 1) % 2].synchronize()
                if on_result is not None:
                    on_result(i - 1, out_host[(i - 1) % len(out_host)])
        if n:
            ps["x_out"][(n - 1) % 2].synchronize()
            if on_result is not None:
                on_result(n - 1, out_host[(n - 1) % len(out_host)])
        main.wait_stream(ps["h2d"])
        main.wait_stream(ps["d2h"])
        return out_host[(n - 1) % len(out_host)] if n else None


class DDPMPipeline:
    config_name = "model_index.json"

    def __init__(self, unet: UNet2DModel, scheduler):
        self.unet = unet
        self.scheduler = scheduler
        self._graph_cache = {}
        self.use_cuda_graph = os.environ.get("DSG_CUDA_GRAPH", "1") != "0"

    # ------------------------------------------------------------------ DiffusionPipeline surface
    @property
    def device(self) -> torch.device:
        return self.unet.device

    def to(self, device=None, dtype=None):
        if device is not None or dtype is not None:
            self.unet = self.unet.to(device=device, dtype=dtype)
        self._graph_cache.clear()
        return self

    def progress_bar(self, iterable):
        from tqdm.auto import tqdm
        cfg = getattr(self, "_progress_bar_config", {})
        return tqdm(iterable, **cfg)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def save_pretrained(self, save_directory: str, safe_serialization: bool = False, variant: Optional[str] = None,
                        **kwargs):
        os.makedirs(save_directory, exist_ok=True)
        index = {"_class_name": self.__class__.__name__, "_diffusers_version": DIFFUSERS_VERSION,
                 "scheduler": ["diffusers", self.scheduler.__class__.__name__],
                 "unet": ["diffusers", self.unet.__class__.__name__]}
        with open(os.path.join(save_directory, self.config_name), "w", encoding="utf-8") as f:
            f.write(json.dumps(index, indent=2, sort_keys=True) + "\n")
        self.unet.save_pretrained(os.path.join(save_directory, "unet"), safe_serialization=safe_serialization,
                                  variant=variant)
        self.scheduler.save_config(os.path.join(save_directory, "scheduler"))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, variant: Optional[str] = None,
                        torch_dtype: Optional[torch.dtype] = None, **kwargs):
        d = pretrained_model_name_or_path
        if not os.path.isdir(d):
            raise EnvironmentError(f"{d} is not a local directory (dsg_b200 loads local checkpoints only)")
        with open(os.path.join(d, cls.config_name), "r", encoding="utf-8") as f:
            index = json.load(f)
        sched_name = index.get("scheduler", ["diffusers", "DDPMScheduler"])[1]
        if sched_name not in _SCHEDULERS:
            raise ValueError(f"scheduler class {sched_name} is not available in dsg_b200")
        sched_cls = _SCHEDULERS[sched_name]
        scheduler = sched_cls.from_config(sched_cls.load_config(os.path.join(d, "scheduler")))
        # upstream uses the variant weights only where the sub-folder actually contains them
        unet_dir = os.path.join(d, "unet")
        use_variant = variant if variant and any(f".{variant}." in f for f in os.listdir(unet_dir)) else None
        unet = UNet2DModel.from_pretrained(d, subfolder="unet", variant=use_variant, torch_dtype=torch_dtype)
        return cls(unet=unet, scheduler=scheduler)

    # ------------------------------------------------------------------ sampling
    def _image_shape(self, batch_size: int) -> Tuple[int, ...]:
        ss = self.unet.config.sample_size
        if isinstance(ss, int):
            return (batch_size, self.unet.config.in_channels, ss, ss)
        return (batch_size, self.unet.config.in_channels, *ss)

    def _postprocess(self, image: torch.Tensor, output_type: str):
        n, c, h, w = image.shape
        lib = _lib.load()
        st = torch.cuda.current_stream(image.device).cuda_stream
        if output_type == "pil":
            u8 = torch.empty((n, h, w, c), dtype=torch.uint8, device=image.device)
            check(lib.dsg_latent_to_image(image.data_ptr(), u8.data_ptr(), None, n, c, h, w, st), "latent_to_image")
            return numpy_to_pil(u8.cpu().numpy())
        f32 = torch.empty((n, h, w, c), dtype=torch.float32, device=image.device)
        check(lib.dsg_latent_to_image(image.data_ptr(), None, f32.data_ptr(), n, c, h, w, st), "latent_to_image")
        return f32.cpu().numpy()

    def _denoise(self, image: torch.Tensor, generator, ddim: bool = False, eta: float = 0.0) -> torch.Tensor:
        dev = image.device
        sched = self.scheduler
        if self.use_cuda_graph:
            key = (tuple(image.shape), ddim, float(eta), sched.num_inference_steps, self.unet._weights_key())
            gs = self._graph_cache.get(key)
            if gs is None:
                self._graph_cache.clear()
                gs = DenoiseSession(self.unet, sched, tuple(image.shape), ddim, eta)
                self._graph_cache[key] = gs
            gs.x.copy_(image)
            for t in self.progress_bar(sched.timesteps):
                ti = int(t)
                gs.set_step(ti)
                if (ddim and eta > 0) or (not ddim and ti > 0):
                    gs.z.copy_(randn_tensor(image.shape, generator=generator, device=dev, dtype=image.dtype))
                gs.graph.replay()
            return gs.x.clone()
        for t in self.progress_bar(sched.timesteps):
            model_output = self.unet(image, t).sample
            if ddim:
                image = sched.step(model_output, t, image, eta=eta, generator=generator).prev_sample
            else:
                image = sched.step(model_output, t, image, generator=generator).prev_sample
        return image

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, num_inference_steps: int = 1000,
                 output_type: Optional[str] = "pil", return_dict: bool = True):
        dev = self.device
        if dev.type != "cuda":
            from .. import testing as _testing
            if _testing.cpu_backend("unet_forward") is None:
                raise DsgError("DDPMPipeline: move the pipeline to a B200 (`.to('cuda')`); no CPU path in dsg_b200")
        shape = self._image_shape(batch_size)
        image = randn_tensor(shape, generator=generator, device=dev)
        self.scheduler.set_timesteps(num_inference_steps)
        if dev.type == "cuda":
            image = self._denoise(image, generator)
            image = self._postprocess(image, output_type)
        else:  # test-suite CPU plumbing only
            for t in self.progress_bar(self.scheduler.timesteps):
                model_output = self.unet(image, t).sample
                image = self.scheduler.step(model_output, t, image, generator=generator).prev_sample
            image = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy()
            if output_type == "pil":
                image = numpy_to_pil(image)
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image)


class DDIMPipeline(DDPMPipeline):
    """``DDIMPipeline.__call__(batch_size, generator, eta, num_inference_steps, use_clipped_model_output, ...)``."""

    def __init__(self, unet: UNet2DModel, scheduler):
        if not isinstance(scheduler, DDIMScheduler):
            scheduler = DDIMScheduler.from_config(scheduler.config) if hasattr(scheduler, "config") else scheduler
        super().__init__(unet, scheduler)

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, eta: float = 0.0, num_inference_steps: int = 50,
                 use_clipped_model_output: Optional[bool] = None, output_type: Optional[str] = "pil",
                 return_dict: bool = True):
        dev = self.device
        if dev.type != "cuda":
            raise DsgError("DDIMPipeline: move the pipeline to a B200 (`.to('cuda')`); no CPU path in dsg_b200")
        shape = self._image_shape(batch_size)
        image = randn_tensor(shape, generator=generator, device=dev)
        self.scheduler.set_timesteps(num_inference_steps)
        image = self._denoise(image, generator, ddim=True, eta=eta)
        image = self._postprocess(image, output_type)
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image)
