"""TrainProgram — the U-Net forward with saved activations plus its hand-written backward, as two flat kernel programs.

Training path of SURVEY.md §8 a17: what torch autograd + cuDNN execute behind ``model(noisy, t)`` /
``accelerator.backward(loss)`` in ``DriveSceneGen/pipeline/training_pipeline.py:84-86``.  Host-side orchestration only —
every arithmetic step is a libdsg_b200 call (include/dsg_b200.h, "training path"):

  forward   the same op sequence as inference (``engine._Program``) but every intermediate a backward op needs
            (GroupNorm+SiLU outputs, conv1 outputs, attention q/k/v and output, time-embedding pre-activations) lives in
            its own buffer instead of a shared temporary.
  backward  per block, in reverse order: bias gradients (column sums), weight gradients (tcgen05 ``dsg_conv_wgrad``),
            data gradients (``dsg_conv`` on the dgrad-packed weights), GroupNorm+SiLU backward (``dsg_gn_bwd``, which
            also adds the shortcut gradient and accumulates into tensors with two consumers), attention backward,
            and finally the time-embedding MLP.  Parameter gradients are written in fp32, torch layout, straight into
            the slices of ONE flat gradient buffer (the single NCCL all-reduce / fused AdamW operate on it).

Activation gradients are fp16 times a power-of-two factor chosen per step from the incoming gradient (``dsg_grad_scale``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch

from ._lib import ConvArgs, WgradArgs, check
from .engine import UNetEngine, _Program, _p


class TrainProgram(_Program):
    regroup = False   # whole-batch launches: the backward op list is built against the forward's whole-batch buffers
    fuse_out_mma = False   # conv_out's weight gradient needs the activated conv_norm_out tensor

    def __init__(self, eng: UNetEngine, batch: int, h: int, w: int, grad_slices: Dict[str, torch.Tensor]):
        """grad_slices: parameter name (upstream state-dict key) -> fp32 view that receives its gradient."""
        self.grads = grad_slices
        self.fuse_gn = False   # the backward needs every normalised tensor (wgrad operand): no in-conv GroupNorm
        self.records: List[dict] = []
        self.bwd_ops: List[Callable[[int], None]] = []
        self.bwd_info: List[Tuple[str, dict]] = []
        self._uid = 0
        self.dout_ptr = C.c_void_p(0)
        self.slot = 0                 # which of the two flat gradient buffers this program writes (hostapi.training)
        self.on_early_ready = None    # hook(program): the "early" slice of the flat gradient buffer is complete
        # CUDA graphs of the forward and of the backward (split where the host hook sits): the ~475 launches of a step
        # leave ~2 ms of gaps between kernels when they are issued one by one.  The first forward + backward of a
        # program run eagerly (lazy module loading, per-kernel attributes), the second forward captures everything.
        self.use_graphs = os.environ.get("DSG_TRAIN_GRAPH", "1") != "0"
        self._eager_fwd = self._eager_bwd = 0
        self._fwd_graph = None
        self._bwd_graphs: List[Tuple[Optional[torch.cuda.CUDAGraph], Optional[Callable]]] = []
        super().__init__(eng, batch, h, w)
        self._build_backward()

    # ------------------------------------------------------------------ forward: every temporary is persistent
    def _build(self):
        self._uid = 0
        self.records = []
        super()._build()

    def _tmp(self, name: str, hw: Tuple[int, int], ch: int) -> torch.Tensor:
        self._uid += 1
        numel = self.b * hw[0] * hw[1] * ch
        return self.eng.arena.get(f"train/{self.b}x{self.h}x{self.w}/{name}{self._uid}", numel, torch.float16)

    def _te_saved(self, half: int, hidden: int) -> Optional[torch.Tensor]:
        self.te_saved = self.eng.arena.get(f"train/{self.b}/te_saved", self.b * (2 * half + 3 * hidden), torch.float32)
        return self.te_saved

    def _attn_lse(self, b: int, heads: int, tokens: int, hd: int) -> Optional[torch.Tensor]:
        if not self.lib.dsg_attention_train_tc_ok(tokens, hd):
            return None
        self._uid += 1
        return self.eng.arena.get(f"train/{self.b}x{self.h}x{self.w}/lse{self._uid}", b * heads * tokens, torch.float32)

    def _record(self, rec: dict):
        self.records.append(rec)

    # ------------------------------------------------------------------ backward construction
    def _btmp(self, name: str, numel: int, dtype=torch.float16) -> torch.Tensor:
        """shared backward scratch (sized for the largest request)."""
        key = f"train/{self.b}x{self.h}x{self.w}/bwd/{name}"
        t = self.eng.arena.bufs.get(key)
        if t is not None and t.numel() >= numel and t.dtype == dtype:
            return t[:numel]
        self._bwd_sizes[(key, dtype)] = max(self._bwd_sizes.get((key, dtype), 0), numel)
        return None

    def _bemit(self, name: str, meta: dict, fn: Callable[[int], None]):
        self.bwd_ops.append(fn)
        self.bwd_info.append((name, meta))

    def _build_backward(self):
        # pass 1 sizes the shared scratch buffers, pass 2 emits the ops (same trick as the forward)
        self._bwd_sizes: Dict[Tuple[str, torch.dtype], int] = {}
        self._sizing = True
        self._gtotal: Dict[int, int] = {}
        self._emit_backward()
        self._gtotal = dict(self._gcount)   # writers per gradient tensor
        for (key, dtype), numel in self._bwd_sizes.items():
            self.eng.arena.get(key, numel, dtype)
        self._sizing = False
        self.bwd_ops, self.bwd_info = [], []
        self._emit_backward()

    # gradient buffer of an activation tensor + whether something has been written to it yet (in backward order)
    def _grad_buf(self, t: torch.Tensor) -> Tuple[torch.Tensor, bool]:
        key = t.data_ptr()
        g = self._gbuf.get(key)
        if g is None:
            self._uid += 1
            g = self.eng.arena.get(f"train/{self.b}x{self.h}x{self.w}/grad{len(self._gbuf)}", t.numel(), torch.float16)
            self._gbuf[key] = g
            self._gwritten[key] = False
            self._gcount[key] = 0
        w = self._gwritten[key]
        self._gwritten[key] = True
        self._gcount[key] += 1
        return g, w

    def _is_last_writer(self, t: Optional[torch.Tensor]) -> bool:
        """call right after _grad_buf(t): True when this is the final write to t's gradient (the writer counts come
        from the sizing pass), i.e. the values stored now are the complete gradient."""
        if t is None or self._sizing:
            return False
        key = t.data_ptr()
        return self._gcount[key] == self._gtotal.get(key, -1)

    def _grad_ready(self, t: torch.Tensor) -> torch.Tensor:
        key = t.data_ptr()
        assert self._gwritten.get(key, False), "backward order error: gradient consumed before it was produced"
        return self._gbuf[key]

    # ---- op helpers (each appends to bwd_ops; in the sizing pass buffers may be None and nothing is emitted)
    def _colsum_to(self, x: torch.Tensor, rows: int, c: int, total, total2=None, scaled=True, split=None):
        """column sums of x [rows][c] -> total (and total2); split = [(tensor, first column, width), ...] sends column
        ranges to different destinations instead (the fused q/k/v bias gradient)."""
        lib = self.lib
        parts = max(1, min(148 * 4, -(-rows // 64)))
        partial = self._btmp(self._uniq("colsum_partial"), parts * c, torch.float32)
        if self._sizing:
            return
        inv = self.inv_scale_ptr if scaled else None
        a1 = (x.data_ptr(), rows, c, partial.data_ptr(), parts)
        if split is None:
            self._colsum_job(partial.data_ptr(), 1, parts, c, None, 0, 0, inv, _p(total), _p(total2))
        else:
            for dst, col0, width in split:
                self._rjobs.append(dict(src=partial.data_ptr() + col0 * 4, n=1, parts=parts, c=width, comps=1,
                                        sample_stride=parts * c, part_stride=c, inv_scale=inv, out0=dst.data_ptr()))
        self._bemit("colsum", {"bytes": rows * c * 2}, lambda st: check(lib.dsg_colsum_h16(*a1, st), "colsum_h16"))

    def _bias_grad(self, t: torch.Tensor, g: torch.Tensor, rows: int, c: int, total, total2=None):
        """bias gradient of the block that produced activation t = column sums of t's complete gradient g: taken from
        the last writer's per-CTA sums when that was a GroupNorm backward, else one more read of g."""
        rec = None if self._sizing else self._osum.get(t.data_ptr())
        if rec is None:
            return self._colsum_to(g, rows, c, total, total2)
        buf, parts = rec
        self._uniq("colsum_partial")   # keep the scratch numbering of the sizing pass (which always took the other path)
        self._colsum_job(buf.data_ptr(), self.b, parts, c, None, 0, 0, self.inv_scale_ptr, _p(total), _p(total2))

    def _wgrad(self, mode, x, dy, hw, cin, cout, grad, ci_total=None, ci_off=0):
        lib = self.lib
        nbytes = max(16, int(lib.dsg_wgrad_workspace_bytes(mode, self.b, hw[0], hw[1], cin, cout)))
        ws = self._btmp("wgrad_ws", (nbytes + 3) // 4, torch.float32)
        if self._sizing:
            return
        a = WgradArgs()
        a.mode, a.n, a.h, a.w, a.cin, a.cout = mode, self.b, hw[0], hw[1], cin, cout
        a.x, a.dy, a.grad = x.data_ptr(), dy.data_ptr(), grad.data_ptr()
        a.ci_total, a.ci_off, a.accumulate = (ci_total or cin), ci_off, 0
        a.inv_scale = self.inv_scale_ptr
        a.workspace, a.workspace_bytes, a.impl = ws.data_ptr(), ws.numel() * 4, 0
        self.keep.append(a)
        ref = C.byref(a)
        opx = {0: hw[0] * hw[1], 1: hw[0] * hw[1] // 4, 2: hw[0] * hw[1] * 4, 3: hw[0] * hw[1]}[mode]
        kk = 1 if mode == 3 else 9
        self._bemit("wgrad", {"flops": 2 * self.b * opx * cout * cin * kk, "mode": mode, "hw": hw, "cin": cin,
                              "cout": cout}, lambda st, r=ref: check(lib.dsg_conv_wgrad(r, st), "conv_wgrad"))

    def _dgrad(self, run_mode, dy, hw_in, cin, cout, wname, out, residual=None, flops_k=None):
        """conv over the gradient dy ([b, hw_in, cin]) with dgrad-packed weights -> out ([.., cout])."""
        if self._sizing:
            return
        eng, lib = self.eng, self.lib
        a = ConvArgs()
        a.mode, a.n, a.h, a.w, a.cin, a.cout = run_mode, self.b, hw_in[0], hw_in[1], cin, cout
        a.x = dy.data_ptr()
        a.wpacked = eng.weights[wname].data_ptr()
        a.residual = _p(residual)
        a.out = out.data_ptr()
        a.impl = eng.conv_impl
        self.keep.append(a)
        ref = C.byref(a)
        self._bemit("dgrad", {"flops": flops_k, "mode": run_mode, "hw": hw_in, "cin": cin, "cout": cout},
                    lambda st, r=ref: check(lib.dsg_conv(r, st), f"dgrad {wname}"))

    def _gn_bwd(self, dy, x1, c1, st1, x2, c2, st2, gname, bname, act, hw, addend, dx1, acc1, dx2, acc2,
                g_gamma, g_beta, colsum_to=None, last1=False, last2=False):
        """colsum_to = (per_n tensor, stride, offset, total tensor) or None.  last1 / last2: this call completes the
        gradient of x1 / x2 — it then also leaves that tensor's column sums (its producer's bias gradient)."""
        eng, lib, b = self.eng, self.lib, self.b
        c, npx = c1 + c2, hw[0] * hw[1]
        wave = int(os.environ.get("DSG_GN_BWD_WAVE", 148 * 2))   # ONE wave of the 2-CTA-per-SM kernels (measured: 2 waves 6-18 % slower)
        chunks = max(1, min(64, wave // b, -(-npx // 64)))
        partial = self._btmp(self._uniq("gn_partial"), b * (chunks + 1) * c * 2 + b * c * 4, torch.float32)
        parts = max(1, min(wave // b, -(-npx // 32))) if (colsum_to or last1 or last2) else 0
        colsum = self._btmp(self._uniq("gn_colsum"), max(1, b * parts * c), torch.float32) if colsum_to else None
        if self._sizing:
            return
        osum = [None, None]
        for i, (last, xs, cw) in enumerate(((last1, x1, c1), (last2, x2, c2))):
            if last:
                buf = eng.arena.get(f"train/{self.b}x{self.h}x{self.w}/osum{len(self._osum)}", b * parts * cw,
                                    torch.float32)
                self._osum[xs.data_ptr()] = (buf, parts)
                osum[i] = buf
        g, bt = eng.weights[gname], eng.weights[bname]
        a1 = (dy.data_ptr(), x1.data_ptr(), c1, st1.data_ptr(), _p(x2), c2, _p(st2), g.data_ptr(), bt.data_ptr(), eng.eps,
              act, partial.data_ptr(), chunks, _p(addend), dx1.data_ptr(), int(acc1), _p(dx2), int(acc2), _p(colsum),
              _p(osum[0]), _p(osum[1]), parts, b, npx, eng.groups)
        # d gamma / d beta: per-sample (sum g, sum g * xh) rows left in slot `chunks` of every sample by the apply kernel
        self._rjobs.append(dict(src=partial.data_ptr() + chunks * c * 2 * 4, n=b, parts=1, c=c, comps=2,
                                sample_stride=(chunks + 1) * c * 2, part_stride=0, inv_scale=self.inv_scale_ptr,
                                out0=g_beta.data_ptr(), out1=g_gamma.data_ptr()))
        if colsum_to:
            per_n, stride, off, total = colsum_to
            self._colsum_job(colsum.data_ptr(), b, parts, c, per_n.data_ptr(), stride, off, self.inv_scale_ptr,
                             _p(total), None)
        nbytes = b * npx * c * 2
        # executed: pass 1 reads x, dy; pass 2 reads x, dy (+ the shortcut addend) and writes dx.  Algorithmic: read x,
        # read dy, write dx.
        self._bemit("gn_bwd", {"bytes": (5 + (1 if addend is not None else 0)) * nbytes,
                               "bytes_alg": 3 * nbytes},
                    lambda st: check(lib.dsg_gn_bwd(*a1, st), "gn_bwd"))

    # ------------------------------------------------------------------ the backward program
    def _emit_backward(self):
        eng, lib, b = self.eng, self.lib, self.b
        G = self.grads
        self._gbuf: Dict[int, torch.Tensor] = {}
        self._gwritten: Dict[int, bool] = {}
        self._gcount: Dict[int, int] = {}
        self._osum: Dict[int, Tuple[torch.Tensor, int]] = {}   # tensor -> per-CTA column sums of its final gradient
        self.scale = eng.arena.get(f"train/{b}/scale", 2, torch.float32)
        self.scale_ptr = self.scale.data_ptr()
        self.inv_scale_ptr = self.scale.data_ptr() + 4
        self.dtemb = eng.arena.get(f"train/{b}/dtemb", b * eng.proj_total, torch.float32)
        # finalisers (d gamma / d beta, bias / time-embedding column sums) are collected and run as ONE launch at the
        # end; their inputs therefore live in per-call scratch buffers instead of shared ones
        self._rjobs: List[dict] = []
        self._ruid = 0
        self._rlaunch = 0
        for rec in reversed(self.records):
            getattr(self, "_bwd_" + rec["kind"])(rec)
        self._emit_reduce_jobs()
        self._bwd_time_embed()

    def _uniq(self, name: str) -> str:
        self._ruid += 1
        return f"{name}#{self._ruid}"

    def _bwd_boundary(self, rec):
        """Between the mid block and the down path: every gradient of the up / mid / output layers has been produced.
        Their finalisers run now (instead of with everybody else's at the end) and the hook lets the data-parallel
        all-reduce of that slice of the flat gradient buffer start under the rest of the backward pass."""
        self._emit_reduce_jobs()
        if not self._sizing:
            self._bemit("early_grads_ready", {}, lambda st: self.on_early_ready(self) if self.on_early_ready else None)

    def _emit_reduce_jobs(self):
        if self._sizing or not self._rjobs:
            self._rjobs = []
            return
        from ._lib import ReduceJob
        arr = (ReduceJob * len(self._rjobs))()
        blk = 0
        for i, j in enumerate(self._rjobs):
            r = arr[i]
            r.src, r.n, r.parts, r.c, r.comps = j["src"], j["n"], j["parts"], j["c"], j["comps"]
            r.sample_stride, r.part_stride = j["sample_stride"], j["part_stride"]
            r.per_n, r.per_n_stride, r.per_n_off = j.get("per_n"), j.get("per_n_stride", 0), j.get("per_n_off", 0)
            r.inv_scale, r.out0, r.out0b, r.out1 = j.get("inv_scale"), j.get("out0"), j.get("out0b"), j.get("out1")
            r.block_begin = blk
            blk += (j["c"] + 31) // 32
        raw = bytes(arr)
        self._rlaunch += 1
        dev = self.eng.arena.get(f"train/{self.b}x{self.h}x{self.w}/reduce_jobs{self._rlaunch}", len(raw), torch.uint8)
        dev.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        lib, njobs, total = self.lib, len(self._rjobs), blk
        ptr = dev.data_ptr()
        self._bemit("colsum_finalize", {"jobs": njobs},
                    lambda st: check(lib.dsg_reduce_rows_batched(ptr, njobs, total, st), "reduce_rows_batched"))
        self._rjobs = []

    def _colsum_job(self, src: int, n: int, parts: int, c: int, per_n, stride, off, inv, total, total2):
        self._rjobs.append(dict(src=src, n=n, parts=parts, c=c, comps=1, sample_stride=parts * c, part_stride=c,
                                per_n=per_n, per_n_stride=stride, per_n_off=off, inv_scale=inv, out0=total, out0b=total2))

    def _bwd_out(self, rec):
        """conv_norm_out + SiLU + conv_out; entry point of the backward (sets the gradient scale)."""
        eng, lib, b = self.eng, self.lib, self.b
        W, G = eng.weights, self.grads
        x, hw, act, c0 = rec["x"], rec["hw"], rec["act"], rec["c0"]
        npx = hw[0] * hw[1]
        numel = b * self.cout * npx
        amax_parts = 148 * 4
        amax_partial = self._btmp("amax_partial", amax_parts, torch.float32)
        wt = self._btmp("conv_out_wt", c0 * self.cout * 9, torch.float32)
        zero_b = self._btmp("zero_bias", c0, torch.float32)
        dact = self._btmp("dact", b * npx * c0)
        sw_parts = min(b * hw[0], 148 * 4)
        sw_partial = self._btmp("small_wgrad_partial", (sw_parts + 1) * (self.cout * 9 * c0 + self.cout), torch.float32)
        gx, acc = self._grad_buf(x)
        last = self._is_last_writer(x)
        if not self._sizing:
            zero_b.zero_()
            w_out = W["conv_out.w"]

            def run(st):
                dout = self.dout_ptr
                check(lib.dsg_grad_scale(dout, numel, amax_partial.data_ptr(), amax_parts, self.scale_ptr, st),
                      "grad_scale")
                check(lib.dsg_conv_out_dgrad_weight(w_out.data_ptr(), self.cout, c0, None, wt.data_ptr(), st),
                      "conv_out_dgrad_weight")
                check(lib.dsg_conv_in_scaled(dout, self.scale_ptr, wt.data_ptr(), zero_b.data_ptr(), dact.data_ptr(), b,
                                             self.cout, hw[0], hw[1], c0, st), "conv_out dgrad")
                check(lib.dsg_small_wgrad(act.data_ptr(), dout, b, hw[0], hw[1], c0, self.cout, 1,
                                          sw_partial.data_ptr(), sw_parts, None, G["conv_out.weight"].data_ptr(),
                                          G["conv_out.bias"].data_ptr(), st), "conv_out wgrad")
            self._bemit("conv_out_bwd", {"bytes": b * npx * (self.cout * 4 * 2 + c0 * 2 * 2)}, run)
        self._gn_bwd(dact, x, c0, rec["st1"], None, 0, None, "norm_out.g", "norm_out.b", 1, hw, None, gx, acc, None, 0,
                     G["conv_norm_out.weight"], G["conv_norm_out.bias"], last1=last)

    def _bwd_in(self, rec):
        """conv_in: weight / bias gradient only (the input image needs no gradient)."""
        lib, b = self.lib, self.b
        G = self.grads
        out, hw, c0 = rec["out"], rec["hw"], rec["c0"]
        g = self._grad_ready(out)
        self._bias_grad(out, g, b * hw[0] * hw[1], c0, G["conv_in.bias"])
        sw_parts = min(b * hw[0], 148 * 4)
        sw_partial = self._btmp("small_wgrad_partial", (sw_parts + 1) * (self.cin * 9 * c0 + self.cin), torch.float32)
        if self._sizing:
            return

        def run(st):
            check(lib.dsg_small_wgrad(g.data_ptr(), self.in_ptr, b, hw[0], hw[1], c0, self.cin, 0, sw_partial.data_ptr(),
                                      sw_parts, self.inv_scale_ptr, G["conv_in.weight"].data_ptr(), None, st),
                  "conv_in wgrad")
        self._bemit("conv_in_bwd", {"bytes": b * hw[0] * hw[1] * (c0 * 2 + self.cin * 4)}, run)

    def _bwd_resnet(self, rec):
        eng, b = self.eng, self.b
        G = self.grads
        r, x1, x2, hw, out = rec["r"], rec["x1"], rec["x2"], rec["hw"], rec["out"]
        c1, c2, co, pre = r["cin"], r["cskip"], r["cout"], r["prefix"]
        c, npx = c1 + c2, hw[0] * hw[1]
        g_out = self._grad_ready(out)
        # conv2 (+ shortcut): bias, weights
        self._bias_grad(out, g_out, b * npx, co, G[f"{pre}.conv2.bias"],
                        G[f"{pre}.conv_shortcut.bias"] if r["has_sc"] else None)
        self._wgrad(0, rec["a2"], g_out, hw, co, co, G[f"{pre}.conv2.weight"])
        if r["has_sc"]:
            self._wgrad(3, x1, g_out, hw, c1, co, G[f"{pre}.conv_shortcut.weight"], ci_total=c, ci_off=0)
            if x2 is not None:
                self._wgrad(3, x2, g_out, hw, c2, co, G[f"{pre}.conv_shortcut.weight"], ci_total=c, ci_off=c1)
        # conv2 data gradient -> GroupNorm2 + SiLU backward -> dh (+ time-embedding / conv1 bias gradients)
        dact = self._btmp("dact", b * npx * max(co, c))
        dh = self._btmp("dh", b * npx * co)
        self._dgrad(0, g_out, hw, co, co, f"{pre}.conv2.dg", dact, flops_k=2 * b * npx * co * co * 9)
        self._gn_bwd(dact, rec["h"], co, rec["h_stats"], None, 0, None, f"{pre}.norm2.g", f"{pre}.norm2.b", 1, hw, None,
                     dh, 0, None, 0, G[f"{pre}.norm2.weight"], G[f"{pre}.norm2.bias"],
                     colsum_to=(self.dtemb, eng.proj_total, r["temb_off"], G[f"{pre}.conv1.bias"]))
        # conv1: weights, data gradient
        self._wgrad(0, rec["a1"], dh, hw, c, co, G[f"{pre}.conv1.weight"])
        self._dgrad(0, dh, hw, co, c, f"{pre}.conv1.dg", dact, flops_k=2 * b * npx * co * c * 9)
        # shortcut gradient w.r.t. cat(x1, x2): 1x1 conv transpose, or the identity
        if r["has_sc"]:
            dsc = self._btmp("dsc", b * npx * c)
            self._dgrad(3, g_out, hw, co, c, f"{pre}.sc.dg", dsc, flops_k=2 * b * npx * co * c)
            addend = dsc
        else:
            assert x2 is None and c1 == co
            addend = g_out
        gx1, acc1 = self._grad_buf(x1)
        last1 = self._is_last_writer(x1)
        gx2, acc2 = self._grad_buf(x2) if x2 is not None else (None, False)
        last2 = self._is_last_writer(x2)
        self._gn_bwd(dact, x1, c1, rec["st1"], x2, c2, rec["st2"], f"{pre}.norm1.g", f"{pre}.norm1.b", 1, hw, addend,
                     gx1, acc1, gx2, acc2, G[f"{pre}.norm1.weight"], G[f"{pre}.norm1.bias"], last1=last1, last2=last2)

    def _bwd_attn(self, rec):
        eng, lib, b = self.eng, self.lib, self.b
        G = self.grads
        a, x, hw, out = rec["a"], rec["x"], rec["hw"], rec["out"]
        ch, pre, hd = a["ch"], a["prefix"], a["head_dim"]
        npx = hw[0] * hw[1]
        g_out = self._grad_ready(out)
        # to_out: bias, weight, data gradient
        self._bias_grad(out, g_out, b * npx, ch, G[f"{pre}.to_out.0.bias"])
        self._wgrad(3, rec["o"], g_out, hw, ch, ch, G[f"{pre}.to_out.0.weight"])
        do = self._btmp("attn_do", b * npx * ch)
        self._dgrad(3, g_out, hw, ch, ch, f"{pre}.out.dg", do, flops_k=2 * b * npx * ch * ch)
        # attention core
        dqkv = self._btmp("attn_dqkv", b * npx * 3 * ch)
        heads = ch // hd
        aws = self._btmp("attn_ws", 2 * b * heads * npx, torch.float32)
        if not self._sizing:
            args = (rec["qkv"].data_ptr(), rec["o"].data_ptr(), do.data_ptr(), dqkv.data_ptr(), aws.data_ptr(),
                    _p(rec["lse"]), b, npx, heads, hd)
            self._bemit("attention_bwd", {"flops": 10 * b * npx * npx * ch},
                        lambda st: check(lib.dsg_attention_bwd(*args, st), "attention_bwd"))
        # q/k/v projections: the fused [3c] gradient is scattered to the three parameter pairs
        qkv_w = self._btmp("attn_qkv_w", 3 * ch * ch, torch.float32)
        self._colsum_to(dqkv, b * npx, 3 * ch, None,
                        split=[(G[f"{pre}.to_{n}.bias"], i * ch, ch) for i, n in enumerate("qkv")])
        self._wgrad(3, rec["act"], dqkv, hw, ch, 3 * ch, qkv_w)
        if not self._sizing:
            dst = [G[f"{pre}.to_{n}.weight"] for n in "qkv"]

            def scatter(st):
                for i, gw in enumerate(dst):
                    gw.view(-1).copy_(qkv_w[i * ch * ch:(i + 1) * ch * ch])
            self._bemit("qkv_scatter", {}, scatter)
        da = self._btmp("dact", b * npx * ch)
        self._dgrad(3, dqkv, hw, 3 * ch, ch, f"{pre}.qkv.dg", da, flops_k=2 * b * npx * 3 * ch * ch)
        gx, acc = self._grad_buf(x)
        last = self._is_last_writer(x)
        self._gn_bwd(da, x, ch, rec["st1"], None, 0, None, f"{pre}.gn.g", f"{pre}.gn.b", 0, hw, g_out, gx, acc, None, 0,
                     G[f"{pre}.group_norm.weight"], G[f"{pre}.group_norm.bias"], last1=last)

    def _bwd_down(self, rec):
        """Downsample2D conv (3x3 stride 2)."""
        b = self.b
        G = self.grads
        pre, x, hw, out, ch = rec["prefix"], rec["x"], rec["hw"], rec["out"], rec["ch"]
        ohw = (hw[0] // 2, hw[1] // 2)
        g_out = self._grad_ready(out)
        self._bias_grad(out, g_out, b * ohw[0] * ohw[1], ch, G[pre + ".bias"])
        self._wgrad(1, x, g_out, hw, ch, ch, G[pre + ".weight"])
        gx, acc = self._grad_buf(x)
        # transpose of the stride-2 conv = a four-phase sub-pixel conv over the low-resolution gradient
        self._dgrad(2, g_out, ohw, ch, ch, pre + ".dg", gx, residual=gx if acc else None,
                    flops_k=2 * b * ohw[0] * ohw[1] * ch * ch * 9)

    def _bwd_up(self, rec):
        """Upsample2D (nearest 2x + 3x3 conv)."""
        b = self.b
        G = self.grads
        pre, x, hw, out, ch = rec["prefix"], rec["x"], rec["hw"], rec["out"], rec["ch"]
        ohw = (hw[0] * 2, hw[1] * 2)
        g_out = self._grad_ready(out)
        self._bias_grad(out, g_out, b * ohw[0] * ohw[1], ch, G[pre + ".bias"])
        self._wgrad(2, x, g_out, hw, ch, ch, G[pre + ".weight"])
        gx, acc = self._grad_buf(x)
        self._dgrad(4, g_out, ohw, ch, ch, pre + ".dg", gx, residual=gx if acc else None,
                    flops_k=2 * b * ohw[0] * ohw[1] * ch * ch * 9)

    def _bwd_time_embed(self):
        """Timesteps -> linear_1 -> SiLU -> linear_2 -> SiLU -> every ResnetBlock's time_emb_proj."""
        eng, lib, b = self.eng, self.lib, self.b
        G, W = self.grads, eng.weights
        hid, P = eng.temb_hidden, eng.proj_total
        in_dim = eng.time_dim
        d_pre2 = self._btmp("te_dpre2", b * hid, torch.float32)
        d_pre1 = self._btmp("te_dpre1", b * hid, torch.float32)
        if self._sizing:
            return
        sv = self.te_saved
        e_ptr = sv.data_ptr()
        pre1_ptr = e_ptr + 4 * b * in_dim
        h1s_ptr = pre1_ptr + 4 * b * hid
        pre2_ptr = h1s_ptr + 4 * b * hid
        inv = self.inv_scale_ptr
        dt = self.dtemb.data_ptr()
        semb = self.emb_ws.data_ptr()
        proj = [(r["temb_off"], r["cout"], G[r["prefix"] + ".time_emb_proj.weight"].data_ptr(),
                 G[r["prefix"] + ".time_emb_proj.bias"].data_ptr()) for r in eng.resnets]
        w2 = W["te.linear_2.w_oi"].data_ptr()
        wp = W["te.proj.w"].data_ptr()
        g1w, g1b = G["time_embedding.linear_1.weight"].data_ptr(), G["time_embedding.linear_1.bias"].data_ptr()
        g2w, g2b = G["time_embedding.linear_2.weight"].data_ptr(), G["time_embedding.linear_2.bias"].data_ptr()

        def run(st):
            for off, co, gw, gb in proj:
                check(lib.dsg_lin_wgrad_small(dt, P, off, semb, hid, co, hid, b, inv, gw, gb, st), "time_emb_proj wgrad")
            check(lib.dsg_lin_dgrad_small(dt, P, 0, wp, P, hid, pre2_ptr, d_pre2.data_ptr(), b, st), "temb dgrad 2")
            check(lib.dsg_lin_wgrad_small(d_pre2.data_ptr(), hid, 0, h1s_ptr, hid, hid, hid, b, inv, g2w, g2b, st),
                  "linear_2 wgrad")
            check(lib.dsg_lin_dgrad_small(d_pre2.data_ptr(), hid, 0, w2, hid, hid, pre1_ptr, d_pre1.data_ptr(), b, st),
                  "temb dgrad 1")
            check(lib.dsg_lin_wgrad_small(d_pre1.data_ptr(), hid, 0, e_ptr, in_dim, hid, in_dim, b, inv, g1w, g1b, st),
                  "linear_1 wgrad")
        self._bemit("time_embed_bwd", {}, run)

    # ------------------------------------------------------------------ execution
    def _capture(self, dev):
        """Static input / output buffers + one graph of the forward + one graph per hook-free stretch of the backward."""
        a, key = self.eng.arena, f"train/{self.b}x{self.h}x{self.w}/graph_io"
        n_in, n_out = self.b * self.cin * self.h * self.w, self.b * self.cout * self.h * self.w
        self.g_in = a.get(key + "/in", n_in, torch.float32).view(self.b, self.cin, self.h, self.w)
        self.g_t = a.get(key + "/t", self.b, torch.float32)
        self.g_out = a.get(key + "/out", n_out, torch.float32).view(self.b, self.cout, self.h, self.w)
        self.g_dout = a.get(key + "/dout", n_out, torch.float32).view(self.b, self.cout, self.h, self.w)
        self.in_ptr = C.c_void_p(self.g_in.data_ptr())
        self.t_ptr = C.c_void_p(self.g_t.data_ptr())
        self.out_ptr = C.c_void_p(self.g_out.data_ptr())
        self.dout_ptr = C.c_void_p(self.g_dout.data_ptr())
        count = self.lib.dsg_launch_count
        with torch.cuda.device(dev):
            g = torch.cuda.CUDAGraph()
            n0 = count()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                st = torch.cuda.current_stream(dev).cuda_stream
                for op in self.ops:
                    op(st)
            self._fwd_graph, self._fwd_launches = g, int(count() - n0)   # reported at every replay
            pool = g.pool()
            # the backward: host hooks (the data-parallel boundary) cannot live inside a graph: split there
            segs, cur = [], []
            for (name, _), op in zip(self.bwd_info, self.bwd_ops):
                if name == "early_grads_ready":
                    segs.append((cur, op))
                    cur = []
                else:
                    cur.append(op)
            segs.append((cur, None))
            self._bwd_graphs = []
            n0 = count()
            for ops, hook in segs:
                gb = None
                if ops:
                    gb = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gb, pool=pool, capture_error_mode="thread_local"):
                        st = torch.cuda.current_stream(dev).cuda_stream
                        for op in ops:
                            op(st)
                self._bwd_graphs.append((gb, hook))
            self._bwd_launches = int(count() - n0)

    def run(self, sample: torch.Tensor, t_float: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self.use_graphs or self._eager_fwd < 1 or self._eager_bwd < 1:
            self._eager_fwd += 1
            self._fwd_graph = None          # an eager run re-points in_ptr / out_ptr: graphs are rebuilt on demand
            return super().run(sample, t_float, out)
        if self._fwd_graph is None:
            try:
                self._capture(sample.device)
            except Exception as e:   # noqa: BLE001 — e.g. a capture-unsafe call on this stream: keep training, launch by launch
                import warnings
                warnings.warn(f"dsg_b200: CUDA-graph capture of the training step failed ({e}); running launch by launch")
                self.use_graphs, self._fwd_graph, self._bwd_graphs = False, None, []
                return super().run(sample, t_float, out)
        with torch.cuda.device(sample.device):
            self.g_in.copy_(sample)
            self.g_t.copy_(t_float)
            self._fwd_graph.replay()
            self.lib.dsg_count_graph_launches(self._fwd_launches)
            if out is None:
                out = torch.empty_like(self.g_out)
            out.copy_(self.g_out)
        return out

    def backward(self, dout: torch.Tensor, sample: torch.Tensor):
        """dout: fp32 NCHW gradient of the model output; sample: the forward's fp32 NCHW input (conv_in's wgrad)."""
        assert dout.dtype == torch.float32 and dout.is_contiguous() and dout.is_cuda
        if self.use_graphs and self._fwd_graph is not None:
            # the forward that is being differentiated ran from g_in (one forward in flight per shape)
            with torch.cuda.device(dout.device):
                self.g_dout.copy_(dout)
                st = torch.cuda.current_stream(dout.device).cuda_stream
                for g, hook in self._bwd_graphs:
                    if g is not None:
                        g.replay()
                    if hook is not None:
                        hook(st)
                self.lib.dsg_count_graph_launches(self._bwd_launches)
            return
        self._eager_bwd += 1
        self.dout_ptr = C.c_void_p(dout.data_ptr())
        self.in_ptr = C.c_void_p(sample.data_ptr())
        with torch.cuda.device(dout.device):
            st = torch.cuda.current_stream(dout.device).cuda_stream
            for op in self.bwd_ops:
                op(st)

    def run_timed(self, sample: torch.Tensor, t_float: torch.Tensor, out: Optional[torch.Tensor] = None):
        graphs, self.use_graphs = self.use_graphs, False   # per-launch events need the eager path
        try:
            return super().run_timed(sample, t_float, out)
        finally:
            self.use_graphs = graphs

    def backward_timed(self, dout: torch.Tensor, sample: torch.Tensor):
        graphs, self.use_graphs = self.use_graphs, False
        try:
            return self._backward_timed(dout, sample)
        finally:
            self.use_graphs = graphs

    def _backward_timed(self, dout: torch.Tensor, sample: torch.Tensor):
        dev = dout.device
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.bwd_ops) + 1)]
        stream = torch.cuda.current_stream(dev)
        ops = self.bwd_ops
        hooked = []
        for i, op in enumerate(ops):
            def h(st, op=op, i=i):
                op(st)
                evs[i + 1].record(stream)
            hooked.append(h)
        self.bwd_ops = hooked
        try:
            evs[0].record(stream)
            self.backward(dout, sample)
        finally:
            self.bwd_ops = ops
        torch.cuda.synchronize(dev)
        return [(n, m, evs[i].elapsed_time(evs[i + 1])) for i, (n, m) in enumerate(self.bwd_info)]
