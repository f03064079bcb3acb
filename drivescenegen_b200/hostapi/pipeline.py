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
    """Two captured CUDA graphs = one denoising step each (timestep bookkeeping + U-Net forward + scheduler step) on
    static device buffers, ping-ponging the sample between two buffers (graph ``p`` reads ``xb[p]`` / ``zb[p]`` and
    writes ``xb[1-p]``), so a step is exactly one graph launch: no per-step host-side argument updates, no copies.

    ``begin(timesteps)`` uploads the schedule; ``advance(noise)`` runs the next step of it; ``step(t, noise)`` runs one
    step at an explicit timestep.  ``x`` is the current sample, ``z`` the variance-noise buffer the next step reads.
    ``run_from_host`` is the end-to-end form: every step's variance noise comes from (pinned) host memory and every
    step's result lands in (pinned) host memory, with the copies on side streams.
    """

    def __init__(self, unet: UNet2DModel, scheduler, shape, ddim: Optional[bool] = None, eta: float = 0.0):
        self.lib = _lib.load()
        dev = unet.device
        self.dev = dev
        is_ddim = isinstance(scheduler, DDIMScheduler)
        if ddim is not None and bool(ddim) != is_ddim:
            raise ValueError(f"DenoiseSession(ddim={ddim}) does not match the scheduler type {type(scheduler).__name__}")
        self.ddim, self.eta = is_ddim, float(eta)
        with torch.cuda.device(dev):
            f32 = dict(dtype=torch.float32, device=dev)
            i32 = dict(dtype=torch.int32, device=dev)
            self.xb = [torch.zeros(shape, **f32), torch.zeros(shape, **f32)]   # the sample, ping-pong
            self.zb = [torch.zeros(shape, **f32), torch.zeros(shape, **f32)]   # variance noise read by graph p
            self.cur = 0
            self.eps = torch.zeros(shape, **f32)
            self.t_f = torch.zeros(shape[0], **f32)       # timestep value (float) per sample
            self.row = torch.zeros(1, **i32)              # coefficient-table row (= timestep)
            self.max_steps = int(scheduler.config.num_train_timesteps)
            self.schedule = torch.zeros(self.max_steps, **i32)
            self.state = torch.zeros(2, **i32)            # {next step index, number of steps}
            self._state_one = torch.tensor([0, 1], **i32)
            self._sched_host = torch.zeros(self.max_steps, dtype=torch.int32).pin_memory()
            self._state_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            self.table = scheduler.coef_table(dev, eta) if is_ddim else scheduler.coef_table(dev)
            eng = unet.engine()
            prog = eng.program(shape[0], shape[2], shape[3])
            step_fn = self.lib.dsg_ddim_step if is_ddim else self.lib.dsg_ddpm_step
            numel = self.eps.numel()

            def body(p):
                st = torch.cuda.current_stream(dev).cuda_stream
                check(self.lib.dsg_step_advance(self.schedule.data_ptr(), self.state.data_ptr(), self.t_f.data_ptr(),
                                                shape[0], self.row.data_ptr(), st), "step advance")
                prog.run(self.xb[p], self.t_f, self.eps)
                check(step_fn(self.eps.data_ptr(), self.xb[p].data_ptr(), self.zb[p].data_ptr(),
                              self.xb[1 - p].data_ptr(), numel, self.table.data_ptr(), self.row.data_ptr(), 0, st),
                      "scheduler step")

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self.state.copy_(self._state_one)
                body(0)  # warm-up outside capture (function attributes, lazy module loading)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graphs = []
            for p in (0, 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self.graphs[0].pool() if self.graphs else None):
                    body(p)
                self.graphs.append(g)
            self.xb[0].zero_(); self.xb[1].zero_()
        self.launches_per_step = prog.n_launches + 2   # + step_advance + scheduler step
        self._pipe = None
        self._hz = None

    # ------------------------------------------------------------------ state
    @property
    def x(self) -> torch.Tensor:
        return self.xb[self.cur]

    @property
    def z(self) -> torch.Tensor:
        return self.zb[self.cur]

    def load(self, sample: torch.Tensor):
        self.x.copy_(sample)

    def begin(self, timesteps):
        """Upload the timestep schedule the following ``advance`` calls walk through."""
        ts = [int(t) for t in timesteps]
        n = len(ts)
        if not 1 <= n <= self.max_steps:
            raise ValueError(f"a schedule needs 1..{self.max_steps} timesteps, got {n}")
        main = torch.cuda.current_stream(self.dev)
        main.synchronize()                       # the pinned staging rows may still be in flight from the last begin()
        self._sched_host[:n] = torch.tensor(ts, dtype=torch.int32)
        self._state_host[0], self._state_host[1] = 0, n
        with torch.cuda.device(self.dev):
            self.schedule[:n].copy_(self._sched_host[:n], non_blocking=True)
            self.state.copy_(self._state_host, non_blocking=True)

    def advance(self, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Run the next step of the schedule (one graph launch); ``noise`` (if given) is copied into ``z`` first."""
        with torch.cuda.device(self.dev):
            if noise is not None:
                self.z.copy_(noise, non_blocking=True)
            self.graphs[self.cur].replay()
        self.cur ^= 1
        return self.x

    def step(self, t: int, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One step at an explicit timestep (two tiny device-side updates, then the graph)."""
        with torch.cuda.device(self.dev):
            self.schedule[0:1].fill_(int(t))
            self.state.copy_(self._state_one)
        return self.advance(noise)

    def step_from_host(self, t: int, noise_host: torch.Tensor, out_host: torch.Tensor):
        self.step(t, noise_host)
        out_host.copy_(self.x, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return out_host

    # ------------------------------------------------------------------ host-generator noise (upstream RNG rule)
    def advance_with_host_generator(self, generator: torch.Generator) -> torch.Tensor:
        """The step's variance noise drawn by ``torch.randn`` with the caller's CPU generator (upstream ``randn_tensor``:
        CPU generator -> draw on the CPU, copy), but straight into pinned memory and copied on a side stream into the
        noise buffer the step after the running one reads: the draw and the copy of step i+1 overlap the graph of step i.
        """
        if self._hz is None:
            with torch.cuda.device(self.dev):
                self._hz = {"buf": [torch.empty(self.eps.shape, dtype=torch.float32).pin_memory() for _ in range(3)],
                            "done": [None, None, None], "i": 0, "stream": torch.cuda.Stream(device=self.dev),
                            "z_in": [torch.cuda.Event(), torch.cuda.Event()],
                            "z_free": [None, None]}
        hz = self._hz
        main = torch.cuda.current_stream(self.dev)
        i, p = hz["i"], self.cur
        slot = hz["buf"][i]
        if hz["done"][i] is not None:
            hz["done"][i].synchronize()           # the copy out of this pinned slot three steps ago
        torch.randn(slot.shape, generator=generator, dtype=torch.float32, out=slot)
        with torch.cuda.device(self.dev), torch.cuda.stream(hz["stream"]):
            if hz["z_free"][p] is not None:
                hz["stream"].wait_event(hz["z_free"][p])   # the step that last read zb[p] has finished
            else:
                hz["stream"].wait_stream(main)
            self.zb[p].copy_(slot, non_blocking=True)
            hz["z_in"][p].record(hz["stream"])
            done = torch.cuda.Event()
            done.record(hz["stream"])
            hz["done"][i] = done
        hz["i"] = (i + 1) % len(hz["buf"])
        main.wait_event(hz["z_in"][p])
        out = self.advance()
        free = torch.cuda.Event()
        free.record(main)
        hz["z_free"][p] = free
        return out

    # ------------------------------------------------------------------ end-to-end loop
    def _pipe_state(self):
        if self._pipe is None:
            with torch.cuda.device(self.dev):
                ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]
                self._pipe = {"h2d": torch.cuda.Stream(device=self.dev), "d2h": torch.cuda.Stream(device=self.dev),
                              "z_in": ev(), "z_free": ev(), "x_ready": ev(), "x_out": ev()}
        return self._pipe

    def run_from_host(self, timesteps, noise_host, out_host, on_result=None):
        """End-to-end loop with the copies taken off the critical path.

        Step ``i`` takes its variance noise from the pinned host tensor ``noise_host[i % len(noise_host)]`` and lands its
        new sample in the pinned host tensor ``out_host[i % len(out_host)]`` (``len(out_host) >= 2``).  The host->device
        copy of step ``i+1`` (into the noise buffer of the other graph) and the device->host copy of step ``i-1`` (from
        the sample buffer step ``i`` only reads) run on their own streams while the graph of step ``i`` computes; the
        host waits for result ``i-1`` after it has queued step ``i`` and then calls ``on_result(i-1, tensor)``.  Every
        step's input still crosses PCIe and every step's result is still read on the host before the call returns.
        """
        if len(out_host) < 2:
            raise ValueError("run_from_host needs at least two host output buffers")
        n = len(timesteps)
        if n == 0:
            return None
        ps = self._pipe_state()
        main = torch.cuda.current_stream(self.dev)
        nz = len(noise_host)
        self.begin(timesteps)
        used = [False, False]         # has graph p run in this call (its z_free / x_out events are valid)?
        p0 = self.cur

        def h2d(i):
            p = (p0 + i) % 2
            with torch.cuda.device(self.dev), torch.cuda.stream(ps["h2d"]):
                if used[p]:
                    ps["h2d"].wait_event(ps["z_free"][p])
                else:
                    ps["h2d"].wait_stream(main)
                self.zb[p].copy_(noise_host[i % nz], non_blocking=True)
                ps["z_in"][p].record(ps["h2d"])

        h2d(0)
        for i in range(n):
            p = (p0 + i) % 2
            if i + 1 < n:
                h2d(i + 1)
            main.wait_event(ps["z_in"][p])
            if used[1 - p]:
                main.wait_event(ps["x_out"][1 - p])   # graph p writes xb[1-p]: its copy-out two steps ago must be done
            self.advance()
            ps["z_free"][p].record(main)
            ps["x_ready"][1 - p].record(main)
            used[p] = True
            with torch.cuda.device(self.dev), torch.cuda.stream(ps["d2h"]):
                ps["d2h"].wait_event(ps["x_ready"][1 - p])
                out_host[i % len(out_host)].copy_(self.xb[1 - p], non_blocking=True)
                ps["x_out"][1 - p].record(ps["d2h"])
            if i >= 1:
                ps["x_out"][p].synchronize()          # result i-1 was copied out of xb[p]
                if on_result is not None:
                    on_result(i - 1, out_host[(i - 1) % len(out_host)])
        ps["x_out"][(p0 + n) % 2].synchronize()
        if on_result is not None:
            on_result(n - 1, out_host[(n - 1) % len(out_host)])
        main.wait_stream(ps["h2d"])
        main.wait_stream(ps["d2h"])
        return out_host[(n - 1) % len(out_host)]


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

    def _denoise(self, image: torch.Tensor, generator, eta: float = 0.0) -> torch.Tensor:
        """The sampling loop.  Which step formula runs follows the TYPE of ``self.scheduler`` (upstream calls
        ``self.scheduler.step`` whatever it is: ``DDPMPipeline(unet, DDIMScheduler())`` is legal and means DDIM, eta 0)."""
        dev = image.device
        sched = self.scheduler
        ddim = isinstance(sched, DDIMScheduler)
        if not ddim and eta:
            raise ValueError("eta is a DDIM parameter; this pipeline holds a DDPMScheduler")
        if self.use_cuda_graph:
            key = (tuple(image.shape), ddim, float(eta), sched.num_inference_steps, id(sched),
                   self.unet._weights_key())
            gs = self._graph_cache.get(key)
            if gs is None:
                self._graph_cache.clear()
                gs = DenoiseSession(self.unet, sched, tuple(image.shape), eta=eta)
                self._graph_cache[key] = gs
            gs.load(image)
            timesteps = [int(t) for t in sched.timesteps]
            gs.begin(timesteps)
            host_rng = isinstance(generator, torch.Generator) and generator.device.type == "cpu"
            for ti in self.progress_bar(timesteps):
                if (ddim and eta > 0) or (not ddim and ti > 0):
                    if host_rng:
                        gs.advance_with_host_generator(generator)
                        continue
                    if generator is None or isinstance(generator, torch.Generator):
                        torch.randn(image.shape, generator=generator, device=dev, dtype=torch.float32, out=gs.z)
                    else:
                        gs.z.copy_(randn_tensor(image.shape, generator=generator, device=dev, dtype=image.dtype))
                gs.advance()
            return self._finite(gs.x.clone())
        for t in self.progress_bar(sched.timesteps):
            model_output = self.unet(image, t).sample
            if ddim:
                image = sched.step(model_output, t, image, eta=eta, generator=generator).prev_sample
            else:
                image = sched.step(model_output, t, image, generator=generator).prev_sample
        return self._finite(image)

    @staticmethod
    def _finite(image: torch.Tensor) -> torch.Tensor:
        """Activations live in HBM as fp16: a badly scaled checkpoint can overflow them.  One reduction over the final
        latent at the end of the run (the only host read of the loop) turns that into an error instead of NaN images."""
        if not bool(torch.isfinite(image).all()):
            raise DsgError("sampling produced non-finite values: activations left the fp16 range inside the U-Net "
                           "(check the checkpoint's GroupNorm gains / weight scale)")
        return image

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, num_inference_steps: int = 1000,
                 output_type: Optional[str] = "pil", return_dict: bool = True):
        dev = self.device
        if dev.type != "cuda":
            raise DsgError("DDPMPipeline: move the pipeline to a B200 (`.to('cuda')`); no CPU path in dsg_b200")
        shape = self._image_shape(batch_size)
        with torch.cuda.device(dev):
            image = randn_tensor(shape, generator=generator, device=dev)
            self.scheduler.set_timesteps(num_inference_steps)
            image = self._denoise(image, generator)
            image = self._postprocess(image, output_type)
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
        with torch.cuda.device(dev):
            image = randn_tensor(shape, generator=generator, device=dev)
            self.scheduler.set_timesteps(num_inference_steps)
            image = self._denoise(image, generator, eta=eta)
            image = self._postprocess(image, output_type)
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image)
