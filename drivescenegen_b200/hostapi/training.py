"""Training path: ``UNet2DModel.forward`` under autograd (DriveSceneGen/pipeline/training_pipeline.py:84-86).

The whole U-Net is ONE ``torch.autograd.Function``: its forward runs the saved-activation forward program of
``engine_train.TrainProgram`` and its backward runs the hand-written backward program (dgrad / wgrad / GroupNorm /
attention kernels of libdsg_b200), which writes every parameter gradient into slices of a single flat fp32 buffer.
The Function hands those slices to autograd as the gradients of the module's parameters, so ``loss.backward()``,
``GradScaler`` and any torch optimizer see ordinary ``.grad`` tensors — no PyTorch arithmetic re-implementation of the
model exists anywhere in the product.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .._lib import DsgError


EARLY_PREFIXES = ("up_blocks.", "mid_block.", "conv_norm_out.", "conv_out.")


def flat_layout(model: torch.nn.Module):
    """Order of the parameters inside the flat parameter / gradient buffers: [late ..., early ...].

    "early" = parameters whose gradients are complete when the backward pass reaches the boundary between the mid block
    and the down path (up blocks, mid block, conv_norm_out, conv_out — except the ``time_emb_proj`` layers, whose
    gradients come out of the time-embedding backward at the very end).  Keeping them contiguous lets the data-parallel
    all-reduce of that slice (64 % of the buffer) start while the down path's backward is still running.
    Returns ([(name, param)], early_offset in elements)."""
    late, early = [], []
    for n, p in model.named_parameters():
        (early if n.startswith(EARLY_PREFIXES) and ".time_emb_proj." not in n else late).append((n, p))
    return late + early, sum(p.numel() for _, p in late)


class _FlatGrads:
    """Two flat fp32 gradient buffers (ping-pong) with per-parameter views in ``flat_layout`` order.

    autograd keeps the views it is handed as ``.grad`` (no copy when it can steal them); if the caller accumulates
    gradients over several backward passes without zeroing, the next pass must not overwrite memory that ``.grad`` still
    aliases — so consecutive passes alternate between the two buffers unless the gradients were cleared.
    """

    def __init__(self, model: torch.nn.Module):
        self.names: List[str] = []
        self.params: List[torch.nn.Parameter] = []
        order, self.early_offset = flat_layout(model)
        for n, p in order:
            self.names.append(n)
            self.params.append(p)
        dev = self.params[0].device
        self.total = sum(p.numel() for p in self.params)
        self.flat = [torch.zeros(self.total, dtype=torch.float32, device=dev) for _ in range(2)]
        self.last = 0   # buffer the most recent backward wrote
        # persistent views: the backward program takes its destination pointers from these
        self.views: List[Dict[str, torch.Tensor]] = [self.fresh_views(i) for i in range(2)]

    def fresh_views(self, which: int) -> Dict[str, torch.Tensor]:
        f, d, off = self.flat[which], {}, 0
        for n, p in zip(self.names, self.params):
            d[n] = f[off:off + p.numel()].view(p.shape)
            off += p.numel()
        return d

    def pick(self) -> int:
        """index of a buffer no live ``.grad`` aliases."""
        g = self.params[0].grad
        if g is not None and g.untyped_storage().data_ptr() == self.flat[0].untyped_storage().data_ptr():
            return 1
        return 0


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, sample, t_float, *params):
        eng = model.engine(train=True)
        b, _, h, w = sample.shape
        fg = model._flat_grads
        which = fg.pick()
        prog = eng.train_program(b, h, w, fg.views[which], which)
        out = prog.run(sample, t_float)
        prog.slot = which
        ctx.prog, ctx.fg, ctx.which, ctx.model = prog, fg, which, model
        ctx.sample = sample
        prog._live_forward = ctx
        return out

    @staticmethod
    def backward(ctx, dout):
        prog = ctx.prog
        if prog._live_forward is not ctx:
            raise DsgError("dsg_b200: backward through a U-Net forward whose saved activations were overwritten by a "
                           "later forward of the same shape (one forward in flight per shape)")
        prog.backward(dout.contiguous().float(), ctx.sample)
        ctx.fg.last = ctx.which
        # fresh view objects (nothing else references them), so autograd can adopt them as .grad without a copy
        views = ctx.fg.fresh_views(ctx.which)
        grads = tuple(views[n] if p.requires_grad else None for n, p in zip(ctx.fg.names, ctx.fg.params))
        del views
        return (None, None, None) + grads


def unet_forward_with_grad(model, sample: torch.Tensor, timestep):
    if sample.requires_grad:
        raise NotImplementedError("dsg_b200: gradients w.r.t. the input sample are not provided (the reference never "
                                  "asks for them); detach the sample")
    b = sample.shape[0]
    dev = sample.device
    t = timestep
    if not torch.is_tensor(t):
        t = torch.full((b,), float(t), dtype=torch.float32, device=dev)
    else:
        t = t.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(b)
        t = t.contiguous()
    if t.numel() != b:
        raise ValueError("timestep must be a scalar or have one entry per sample")
    if getattr(model, "_flat_grads", None) is None or model._flat_grads.params[0].device != dev:
        model._flat_grads = _FlatGrads(model)
    in_dtype = sample.dtype
    x = sample.detach().float().contiguous()
    out = _UNetFn.apply(model, x, t, *model._flat_grads.params)
    return out if in_dtype == torch.float32 else out.to(in_dtype)
