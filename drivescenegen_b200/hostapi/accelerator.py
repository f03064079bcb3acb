"""``accelerate.Accelerator`` / ``notebook_launcher`` mirror — the only distributed component of the reference.

Reference call sites (DriveSceneGen/pipeline/training_pipeline.py): ctor :48-53, ``is_main_process`` :54,100,
``init_trackers`` :56, ``prepare`` :59, ``is_local_main_process`` :67, ``accumulate`` :82, ``backward`` :86,
``clip_grad_norm_`` :88, ``log`` :96, ``unwrap_model`` :101; ``notebook_launcher(fn, args, num_processes=1)``
DriveSceneGen/scripts/train.py:122.  Semantics restated from accelerate 0.22.0 in SURVEY.md App. B.4.

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE).  Data parallelism is ONE all-reduce per step over a
single flat gradient buffer (NCCL over NVLink/NVSwitch; gloo on CPU for the tests), issued from ``backward``; the mean
(1/world) is folded into the same pass.  No DistributedDataParallel wrapper, so ``unwrap_model`` is the identity.
"""
from __future__ import annotations

import contextlib
import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class AcceleratedOptimizer:
    def __init__(self, optimizer, accelerator: "Accelerator"):
        self.optimizer = optimizer
        self.accelerator = accelerator
        self.step_was_skipped = False

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    @property
    def state(self):
        self.accelerator._export_fused_state(self.optimizer)
        return self.optimizer.state

    def state_dict(self):
        self.accelerator._export_fused_state(self.optimizer)
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd)
        self.accelerator._import_fused_state(self.optimizer)

    def zero_grad(self, set_to_none: Optional[bool] = None):
        if self.accelerator.sync_gradients:
            if set_to_none is None:
                self.optimizer.zero_grad()
            else:
                self.optimizer.zero_grad(set_to_none=set_to_none)

    def step(self, closure=None):
        acc = self.accelerator
        if not acc.sync_gradients:
            return
        if closure is None and acc._fused_adamw_step(self):
            return
        if acc.scaler is not None:
            before = acc.scaler.get_scale()
            acc.scaler.step(self.optimizer, closure) if closure is not None else acc.scaler.step(self.optimizer)
            acc.scaler.update()
            self.step_was_skipped = acc.scaler.get_scale() < before
            # the scaler moved without the fused path seeing it: re-read its scale / growth counter next time
            acc._scale_host = None
            acc._growth_tracker = int(acc.scaler._get_growth_tracker())
        else:
            self.optimizer.step(closure) if closure is not None else self.optimizer.step()
            self.step_was_skipped = False


class AcceleratedScheduler:
    def __init__(self, scheduler, optimizers: List[AcceleratedOptimizer], accelerator: "Accelerator"):
        self.scheduler = scheduler
        self.optimizers = optimizers
        self.accelerator = accelerator

    def step(self, *args, **kwargs):
        acc = self.accelerator
        if not acc.sync_gradients:
            return
        if any(o.step_was_skipped for o in self.optimizers):
            return
        # upstream: when batches are not split across ranks the wrapped scheduler advances num_processes times
        for _ in range(acc.num_processes if not acc.split_batches else 1):
            self.scheduler.step(*args, **kwargs)

    def get_last_lr(self):
        return self.scheduler.get_last_lr()

    def state_dict(self):
        return self.scheduler.state_dict()

    def load_state_dict(self, sd):
        self.scheduler.load_state_dict(sd)

    def __getattr__(self, name):
        return getattr(self.scheduler, name)


class BatchSamplerShard:
    """Rank r of N takes batches r, r+N, r+2N, ... of the wrapped batch sampler — only the INDICES are enumerated on
    every rank; each rank then loads (decodes) its own samples only.  accelerate 0.22.0 semantics for
    ``split_batches=False, even_batches=True``: every rank runs the same number of steps with the
    same batch size; a short final batch and the missing batches of the last round are filled with samples from the
    beginning of the epoch's index stream."""

    def __init__(self, batch_sampler, rank: int, world: int):
        self.batch_sampler, self.rank, self.world = batch_sampler, rank, world
        self.batch_size = getattr(batch_sampler, "batch_size", None)
        self.drop_last = bool(getattr(batch_sampler, "drop_last", False))

    def __len__(self):
        n = len(self.batch_sampler)
        return n // self.world if self.drop_last else (n + self.world - 1) // self.world

    def __iter__(self):
        world, rank, bs = self.world, self.rank, self.batch_size
        head: List[int] = []        # indices of the epoch's first `world` batches: what the tail is padded from
        n_head = 0
        rnd: List[List[int]] = []   # the batches of the round being collected (one per rank)
        for batch in self.batch_sampler:
            batch = list(batch)
            if n_head < world:
                head.extend(batch)
                n_head += 1
            rnd.append(batch)
            if len(rnd) == world and (bs is None or len(batch) == bs):
                yield rnd[rank]
                rnd = []
        if not rnd or self.drop_last or not head:
            return
        if bs is None:
            bs = len(rnd[0])
        pool = head
        while len(pool) < world * bs:
            pool = pool + pool
        pos = 0
        if len(rnd[-1]) < bs:
            need = bs - len(rnd[-1])
            rnd[-1] = rnd[-1] + pool[pos:pos + need]
            pos += need
        while len(rnd) < world:
            rnd.append(pool[pos:pos + bs])
            pos += bs
        yield rnd[rank]


class ShardedDataLoader:
    """The prepared DataLoader: batches moved to the device; with N ranks, rank r sees batches r, r+N, r+2N, ... (per-rank
    batch size unchanged) through a DataLoader rebuilt over ``BatchSamplerShard``, and the shuffle generator is
    synchronised from rank 0 at the start of every epoch (upstream ``synchronize_rng_states(["generator"])``), so the
    ranks' shards stay disjoint whatever each rank did to its global RNG in between (the reference's ``evaluate`` calls
    ``torch.manual_seed`` on rank 0 only, training_pipeline.py:29)."""

    def __init__(self, loader, device, rank: int, world: int):
        self.device, self.rank, self.world = device, rank, world
        self.base = loader
        self.sync_generator = None
        bsamp = getattr(loader, "batch_sampler", None)
        if world > 1 and bsamp is not None and not isinstance(loader.dataset, torch.utils.data.IterableDataset):
            sampler = getattr(bsamp, "sampler", None)
            if isinstance(sampler, torch.utils.data.RandomSampler):
                if sampler.generator is None:
                    sampler.generator = torch.Generator()
                self.sync_generator = sampler.generator
            kw = dict(num_workers=loader.num_workers, collate_fn=loader.collate_fn, pin_memory=loader.pin_memory,
                      timeout=loader.timeout, worker_init_fn=loader.worker_init_fn,
                      multiprocessing_context=loader.multiprocessing_context, generator=loader.generator,
                      persistent_workers=loader.persistent_workers)
            if loader.num_workers > 0:
                kw["prefetch_factor"] = loader.prefetch_factor
            self.loader = torch.utils.data.DataLoader(loader.dataset, batch_sampler=BatchSamplerShard(bsamp, rank, world),
                                                      **kw)
            self.sharded_by_sampler = True
        else:
            self.loader = loader
            self.sharded_by_sampler = False

    def __len__(self):
        if self.sharded_by_sampler or self.world == 1:
            return len(self.loader)
        return (len(self.loader) + self.world - 1) // self.world

    @property
    def dataset(self):
        return self.loader.dataset

    @property
    def batch_size(self):
        return self.base.batch_size

    def _move(self, b):
        if torch.is_tensor(b):
            if self.device.type == "cuda":
                from . import raster
                if raster.is_raster_batch(b, self.base.dataset):
                    # RasterDataset batches cross PCIe as bytes; resample + normalise run on the device
                    src = b if b.is_cuda or b.is_pinned() else b.pin_memory()
                    size = getattr(self.base.dataset, "size", None)
                    return raster.image_to_sample(src.to(self.device, non_blocking=True), channels=min(3, b.shape[3]),
                                                  size=size)
            return b.to(self.device, non_blocking=True)
        if isinstance(b, (list, tuple)):
            return type(b)(self._move(x) for x in b)
        if isinstance(b, dict):
            return {k: self._move(v) for k, v in b.items()}
        return b

    def _sync_rng(self):
        if self.sync_generator is None or not dist.is_initialized():
            return
        state = self.sync_generator.get_state()
        dev = self.device if dist.get_backend() == "nccl" else torch.device("cpu")
        t = state.to(dev)
        dist.broadcast(t, src=0)
        self.sync_generator.set_state(t.cpu())

    def __iter__(self):
        if self.world == 1 or self.sharded_by_sampler:
            self._sync_rng()
            for b in self.loader:
                yield self._move(b)
            return
        # no batch sampler to shard (iterable dataset): every rank walks the stream and keeps its own batches; the tail
        # is padded by wrapping around so that every rank runs the same number of steps
        batches, first = [], []
        for i, b in enumerate(self.loader):
            if len(first) < self.world:
                first.append(b)
            batches.append(b)
            if len(batches) == self.world:
                yield self._move(batches[self.rank])
                batches = []
        if batches:
            while len(batches) < self.world:
                batches.append(first[len(batches) % len(first)])
            yield self._move(batches[self.rank])


class Accelerator:
    def __init__(self, mixed_precision: Optional[str] = None, gradient_accumulation_steps: int = 1,
                 log_with=None, project_dir: Optional[str] = None, split_batches: bool = False, cpu: bool = False,
                 **kwargs):
        self.mixed_precision = (mixed_precision or os.environ.get("ACCELERATE_MIXED_PRECISION", "no")).lower()
        if self.mixed_precision not in ("no", "fp16", "bf16"):
            raise ValueError(f"Unknown mixed_precision mode: {self.mixed_precision}")
        self.gradient_accumulation_steps = int(gradient_accumulation_steps)
        self.log_with = log_with
        self.project_dir = project_dir
        self.split_batches = split_batches
        self.process_index = int(os.environ.get("RANK", "0"))
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
        use_cuda = torch.cuda.is_available() and not cpu
        self.device = torch.device("cuda", self.local_process_index) if use_cuda else torch.device("cpu")
        if use_cuda:
            torch.cuda.set_device(self.device)
        if self.num_processes > 1 and not dist.is_initialized():
            dist.init_process_group(backend="nccl" if use_cuda else "gloo")
        self.scaler = None
        if self.mixed_precision == "fp16" and use_cuda:
            self.scaler = torch.amp.GradScaler("cuda")
        self.sync_gradients = True
        self.step = 0
        self._models: List[torch.nn.Module] = []
        self._optimizers: List[AcceleratedOptimizer] = []
        self._flat_grad: Optional[torch.Tensor] = None
        self._fused_state = {}
        self._ctl = None
        self._early = None
        self._scale_host: Optional[float] = None
        self._growth_tracker = 0
        self.trackers = []
        self._writer = None

    # ------------------------------------------------------------------ process info
    @property
    def is_main_process(self) -> bool:
        return self.process_index == 0

    @property
    def is_local_main_process(self) -> bool:
        return self.local_process_index == 0

    @property
    def use_distributed(self) -> bool:
        return self.num_processes > 1

    def wait_for_everyone(self):
        if self.use_distributed:
            dist.barrier()

    def print(self, *a, **k):
        if self.is_local_main_process:
            print(*a, **k)

    # ------------------------------------------------------------------ trackers
    def init_trackers(self, project_name: str, config: Optional[dict] = None, init_kwargs: dict = {}):
        if not self.is_main_process or self.log_with is None:
            return
        if self.log_with not in ("tensorboard", "all") and "tensorboard" not in str(self.log_with):
            return
        from torch.utils.tensorboard import SummaryWriter
        logdir = os.path.join(self.project_dir or ".", project_name)
        os.makedirs(logdir, exist_ok=True)
        self._writer = SummaryWriter(logdir)

    def log(self, values: dict, step: Optional[int] = None, log_kwargs: dict = {}):
        if self._writer is None:
            return
        for k, v in values.items():
            if isinstance(v, (int, float)):
                self._writer.add_scalar(k, v, global_step=step)
            elif isinstance(v, str):
                self._writer.add_text(k, v, global_step=step)
        self._writer.flush()

    def end_training(self):
        if self._writer is not None:
            self._writer.close()
            self._writer = None

    # ------------------------------------------------------------------ prepare
    def prepare(self, *args):
        out = []
        for obj in args:
            if isinstance(obj, torch.nn.Module):
                out.append(self.prepare_model(obj))
            elif isinstance(obj, torch.optim.Optimizer):
                o = AcceleratedOptimizer(obj, self)
                self._optimizers.append(o)
                out.append(o)
            elif isinstance(obj, torch.utils.data.DataLoader):
                out.append(ShardedDataLoader(obj, self.device, self.process_index, self.num_processes))
            elif isinstance(obj, torch.optim.lr_scheduler.LRScheduler):
                out.append(("__sched__", obj))
            else:
                out.append(obj)
        out = [AcceleratedScheduler(o[1], self._optimizers, self) if isinstance(o, tuple) and o and o[0] == "__sched__"
               else o for o in out]
        return out[0] if len(out) == 1 else tuple(out)

    def prepare_model(self, model: torch.nn.Module):
        model = model.to(self.device)
        if self.use_distributed:
            # replicas start identical: broadcast rank 0's parameters/buffers once
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0)
        if self.device.type == "cuda" and hasattr(model, "engine"):
            self._flatten_parameters(model)
        self._models.append(model)
        return model

    # ------------------------------------------------------------------ flat buffers + fused optimizer (CUDA engine)
    @staticmethod
    def _flatten_parameters(model: torch.nn.Module):
        """Re-home every parameter in ONE flat fp32 buffer (same Parameter objects, same values): the fused AdamW
        kernel and the single gradient all-reduce then work on whole buffers instead of 160 small tensors."""
        from .training import flat_layout
        params = [p for _, p in flat_layout(model)[0]]   # the order of the flat GRADIENT buffers (hostapi.training)
        if not params or getattr(model, "_flat_params", None) is not None:
            return
        if any(p.dtype != torch.float32 for p in params):
            return
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
        off = 0
        offsets = {}
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            offsets[id(p)] = off
            off += n
        model._flat_params = flat
        model._flat_offsets = offsets

    def _flat_grad_of(self, model) -> Optional[torch.Tensor]:
        """the flat gradient buffer the engine's last backward wrote, if the parameters' .grad alias it."""
        fg = getattr(model, "_flat_grads", None)
        if fg is None:
            return None
        flat = fg.flat[fg.last]
        first, last = fg.params[0].grad, fg.params[-1].grad
        if first is None or last is None:
            return None
        if first.data_ptr() != flat.data_ptr():
            return None
        if last.data_ptr() != flat.data_ptr() + 4 * (flat.numel() - last.numel()):
            return None
        return flat

    @staticmethod
    def _same_params(a, b) -> bool:
        """the same set of Parameter objects (the flat buffers keep their own order, see hostapi.training.flat_layout)"""
        return len(a) == len(b) and {id(p) for p in a} == {id(p) for p in b}

    def _loss_scale_value(self) -> float:
        if self.scaler is None:
            return 1.0
        if self._scale_host is None:
            self._scale_host = float(self.scaler.get_scale())
        return self._scale_host

    def _fused_grad_norm(self, model, max_norm: float) -> Optional[torch.Tensor]:
        flat_g = self._flat_grad_of(model)
        if flat_g is None:
            return None
        from .. import ops
        ctl = ops.grad_norm(flat_g, 1.0 / self._loss_scale_value(), float(max_norm))
        self._ctl = (model, flat_g.data_ptr(), ctl)
        return ctl

    def _fused_adamw_step(self, wrapped: "AcceleratedOptimizer") -> bool:
        """clip + unscale + AdamW in one pass over the flat buffers (dsg_adamw_step); False = not applicable."""
        opt = wrapped.optimizer
        if type(opt) is not torch.optim.AdamW or len(opt.param_groups) != 1 or len(self._models) != 1:
            return False
        model = self._models[0]
        flat_p = getattr(model, "_flat_params", None)
        grp = opt.param_groups[0]
        if flat_p is None or grp.get("amsgrad") or grp.get("maximize") or grp.get("capturable") \
                or grp.get("differentiable"):
            return False
        fg = getattr(model, "_flat_grads", None)
        if fg is None or not self._same_params(grp["params"], fg.params):
            return False
        if fg.params[0].data_ptr() != flat_p.data_ptr():
            return False
        flat_g = self._flat_grad_of(model)
        if flat_g is None:
            return False
        ctl = None
        if self._ctl is not None and self._ctl[0] is model and self._ctl[1] == flat_g.data_ptr():
            ctl = self._ctl[2]
        if ctl is None:
            ctl = self._fused_grad_norm(model, 0.0)
        self._ctl = None
        st = self._fused_state.get(id(opt))
        if st is None:
            st = {"m": torch.zeros_like(flat_p), "v": torch.zeros_like(flat_p), "step": 0}
            self._fused_state[id(opt)] = st
        from .. import ops
        lr = grp["lr"]
        lr = float(lr.item()) if torch.is_tensor(lr) else float(lr)
        b1, b2 = grp["betas"]
        # The kernel itself skips the update when the gradients are not finite (ctl[2]); it is queued together with
        # the re-pack of the fp16 weights for the next forward BEFORE the one host read of the step, so the GPU keeps
        # working while the host waits for the flag (GradScaler / LR-scheduler bookkeeping needs it).
        ops.adamw_step(flat_p, flat_g, st["m"], st["v"], lr, float(b1), float(b2), float(grp["eps"]),
                       float(grp["weight_decay"]), st["step"] + 1, ctl)
        model._weights_epoch = getattr(model, "_weights_epoch", 0) + 1
        if hasattr(model, "engine") and getattr(model, "_engine", None) is not None and model._engine.train_packs:
            model.engine(train=True)
        skipped = bool(ctl[2].item() != 0.0)
        if not skipped:
            st["step"] += 1
        wrapped.step_was_skipped = skipped
        if self.scaler is not None:
            scale = self._loss_scale_value()
            if skipped:
                scale *= self.scaler.get_backoff_factor()
                self._growth_tracker = 0
            else:
                self._growth_tracker += 1
                if self._growth_tracker >= self.scaler.get_growth_interval():
                    scale *= self.scaler.get_growth_factor()
                    self._growth_tracker = 0
            if scale != self._scale_host:
                self.scaler.update(new_scale=scale)
                self._scale_host = scale
        return True

    # The fused kernel keeps AdamW's moments in two flat buffers.  torch's optimizer state is made to ALIAS them
    # (per-parameter views), so optimizer.state / state_dict() / load_state_dict() — checkpoint and resume — see the
    # real moments and the real bias-correction step.
    def _export_fused_state(self, opt):
        st = self._fused_state.get(id(opt))
        if st is None:
            return
        offsets = self._models[0]._flat_offsets
        for p in opt.param_groups[0]["params"]:
            n, off = p.numel(), offsets[id(p)]
            s = opt.state[p]
            m = s.get("exp_avg")
            if m is None or m.data_ptr() != st["m"].data_ptr() + 4 * off:
                s["exp_avg"] = st["m"][off:off + n].view_as(p)
                s["exp_avg_sq"] = st["v"][off:off + n].view_as(p)
            s["step"] = torch.tensor(float(st["step"]))

    def _import_fused_state(self, opt):
        if len(opt.param_groups) != 1:
            return
        params = opt.param_groups[0]["params"]
        if not params or not all("exp_avg" in opt.state.get(p, {}) for p in params):
            return
        if len(self._models) != 1 or getattr(self._models[0], "_flat_params", None) is None:
            return
        flat_p = self._models[0]._flat_params
        if sum(p.numel() for p in params) != flat_p.numel():
            return
        st = self._fused_state.get(id(opt))
        if st is None:
            st = {"m": torch.zeros_like(flat_p), "v": torch.zeros_like(flat_p), "step": 0}
            self._fused_state[id(opt)] = st
        offsets = self._models[0]._flat_offsets
        for p in params:
            n, off = p.numel(), offsets[id(p)]
            s = opt.state[p]
            st["m"][off:off + n].copy_(s["exp_avg"].reshape(-1))
            st["v"][off:off + n].copy_(s["exp_avg_sq"].reshape(-1))
        st["step"] = int(float(opt.state[params[0]]["step"]))
        self._export_fused_state(opt)

    def unwrap_model(self, model, keep_fp32_wrapper: bool = True):
        return model

    # ------------------------------------------------------------------ training step plumbing
    @contextlib.contextmanager
    def accumulate(self, *models):
        self.step += 1
        self.sync_gradients = (self.step % self.gradient_accumulation_steps) == 0
        yield

    @contextlib.contextmanager
    def autocast(self):
        if self.device.type == "cuda" and self.mixed_precision in ("fp16", "bf16"):
            dt = torch.float16 if self.mixed_precision == "fp16" else torch.bfloat16
            with torch.autocast("cuda", dtype=dt):
                yield
        else:
            yield

    def _reduce_mean(self, t: torch.Tensor, async_op: bool = False):
        """in-place mean over ranks: NCCL averages inside the collective; gloo sums and the division is a second pass"""
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op), False
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op), True

    def _arm_early_allreduce(self):
        """Let the backward pass start the all-reduce of the gradients that are complete first (up / mid / output layers:
        the tail of the flat buffer, hostapi.training.flat_layout) while the down path's backward is still computing.
        NCCL runs it on the process group's own stream, ordered after the point of the backward where it was issued; the
        rest of the buffer follows when the backward returns.  Same single flat buffer, same values: the collective is
        only split in two so that most of it hides under compute."""
        self._early = None
        if not (self.use_distributed and self.sync_gradients and len(self._models) == 1) \
                or os.environ.get("DSG_AR_OVERLAP", "1") == "0":
            return
        model = self._models[0]
        fg = getattr(model, "_flat_grads", None)
        eng = getattr(model, "_engine", None)
        if fg is None or eng is None or any(p.grad is not None for p in fg.params):
            return   # gradients are being accumulated into existing .grad tensors: the flat buffer is not the final word

        def hook(prog):
            flat = fg.flat[prog.slot]
            work, needs_div = self._reduce_mean(flat[fg.early_offset:], async_op=True)
            self._early = (flat.data_ptr(), work, needs_div)

        for prog in eng.train_programs.values():
            prog.on_early_ready = hook

    def _disarm_early_allreduce(self):
        model = self._models[0] if len(self._models) == 1 else None
        eng = getattr(model, "_engine", None) if model is not None else None
        if eng is not None:
            for prog in eng.train_programs.values():
                prog.on_early_ready = None

    def _allreduce_grads(self):
        """ONE flat gradient buffer, averaged over ranks; its tail may already be in flight."""
        if len(self._models) == 1:
            flat = self._flat_grad_of(self._models[0])
            early = getattr(self, "_early", None)
            self._early = None
            if flat is not None:   # the engine's backward already wrote one flat buffer: reduce it in place
                if early is not None and early[0] == flat.data_ptr():
                    off = self._models[0]._flat_grads.early_offset
                    _, div = self._reduce_mean(flat[:off])
                    early[1].wait()
                    if div or early[2]:
                        flat.mul_(1.0 / self.num_processes)
                    return
                _, div = self._reduce_mean(flat)
                if div:
                    flat.mul_(1.0 / self.num_processes)
                return
            if early is not None:
                early[1].wait()   # (cannot happen: the hook only fires when .grad will alias the flat buffer)
        params = [p for m in self._models for p in m.parameters() if p.grad is not None]
        if not params:
            return
        total = sum(p.grad.numel() for p in params)
        if self._flat_grad is None or self._flat_grad.numel() != total or self._flat_grad.device != params[0].device:
            self._flat_grad = torch.empty(total, dtype=torch.float32, device=params[0].grad.device)
        flat = self._flat_grad
        views = []
        off = 0
        for p in params:
            n = p.grad.numel()
            views.append(flat[off:off + n].view_as(p.grad))
            off += n
        torch._foreach_copy_(views, [p.grad for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / self.num_processes)
        torch._foreach_copy_([p.grad for p in params], views)

    def backward(self, loss: torch.Tensor, **kwargs):
        loss = loss / self.gradient_accumulation_steps
        self._arm_early_allreduce()
        try:
            if self.scaler is not None:
                self.scaler.scale(loss).backward(**kwargs)
            else:
                loss.backward(**kwargs)
        finally:
            if self.use_distributed:
                self._disarm_early_allreduce()
        if self.use_distributed and self.sync_gradients:
            self._allreduce_grads()

    def unscale_gradients(self):
        if self.scaler is not None:
            for o in self._optimizers:
                self.scaler.unscale_(o.optimizer)

    def clip_grad_norm_(self, parameters: Iterable[torch.Tensor], max_norm: float, norm_type: float = 2):
        if norm_type == 2 and len(self._models) == 1 and len(self._optimizers) == 1 \
                and type(self._optimizers[0].optimizer) is torch.optim.AdamW \
                and getattr(self._models[0], "_flat_params", None) is not None:
            # CUDA engine: the norm of the flat gradient buffer; unscale + clip are applied inside the fused AdamW
            # kernel (the .grad tensors keep their scaled values)
            plist = list(parameters)
            fg = getattr(self._models[0], "_flat_grads", None)
            if fg is not None and self._same_params(plist, fg.params):
                ctl = self._fused_grad_norm(self._models[0], max_norm)
                if ctl is not None:
                    return ctl[0]
            parameters = plist
        self.unscale_gradients()
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type)

    def gather(self, tensor: torch.Tensor):
        if not self.use_distributed:
            return tensor
        outs = [torch.empty_like(tensor) for _ in range(self.num_processes)]
        dist.all_gather(outs, tensor.contiguous())
        return torch.cat(outs, dim=0)


def notebook_launcher(function, args=(), num_processes=None, mixed_precision="no", use_port="29500", **kwargs):
    """upstream with num_processes=1 (the reference's setting): banner + in-process call.  Multi-GPU runs launch the
    script under torchrun; each rank's ``Accelerator()`` reads RANK / LOCAL_RANK / WORLD_SIZE."""
    if num_processes is not None and num_processes > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        raise NotImplementedError("dsg_b200: launch multi-GPU training with torchrun "
                                  "(python -m torch.distributed.run --nproc-per-node N script.py)")
    dev = "GPU" if torch.cuda.is_available() else "CPU"
    print(f"Launching training on one {dev}." if int(os.environ.get("WORLD_SIZE", "1")) == 1
          else f"Launching training on {os.environ['WORLD_SIZE']} processes.")
    function(*args)
