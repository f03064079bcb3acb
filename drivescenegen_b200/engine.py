"""UNetEngine — schedules the U-Net forward as a flat program of libdsg_b200 calls.

Host-side orchestration only: every arithmetic step of ``UNet2DModel.forward`` (diffusers 0.20.0
``models/unet_2d.py``; reference call sites ``DriveSceneGen/pipeline/training_pipeline.py:84`` and, through
``DDPMPipeline.__call__``, ``DriveSceneGen/scripts/generation.py:14``) runs in the sm_100a kernels behind the C ABI of
``include/dsg_b200.h``.  PyTorch is used for device memory and streams.  The program for one input shape is built once
(static buffers, static pointers), so it can be replayed eagerly or captured in a CUDA graph.

Data layout: activations are NHWC fp16; model input/output are NCHW fp32 (the reference's layout).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvArgs, DsgError, check


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _Arena:
    """Named persistent device buffers; a name requested again with a larger size is re-allocated."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}
        self.retired: List[torch.Tensor] = []  # outgrown buffers stay alive: older programs hold raw pointers

    def get(self, name: str, numel: int, dtype) -> torch.Tensor:
        t = self.bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            if t is not None:
                self.retired.append(t)
            t = torch.empty(max(numel, 1), dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t[:numel]

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in list(self.bufs.values()) + self.retired)


class UNetEngine:
    """Packed weights + per-shape execution programs for one ``UNet2DModel`` configuration."""

    def __init__(self, config: dict, device: torch.device):
        self.lib = _lib.load()
        self.cfg = dict(config)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise DsgError("UNetEngine needs a CUDA device (there is no CPU fallback)")
        boc = list(self.cfg["block_out_channels"])
        for c in boc:
            if c % 64 != 0:
                raise DsgError(f"block_out_channels must be multiples of 64 for the tcgen05 path, got {boc}")
        if self.cfg.get("mid_block_scale_factor", 1) != 1:
            raise DsgError("mid_block_scale_factor != 1 is not supported")
        if self.cfg.get("center_input_sample", False):
            raise DsgError("center_input_sample=True is not supported")
        self.groups = int(self.cfg.get("norm_num_groups", 32))
        self.eps = float(self.cfg.get("norm_eps", 1e-5))
        self.arena = _Arena(self.device)
        self.weights: Dict[str, torch.Tensor] = {}
        self.programs: Dict[Tuple[int, int, int], "_Program"] = {}
        self.packed = False
        self.conv_impl = 0       # 0 = tcgen05 igemm, 1 = plain CUDA cross-check kernel (tests only)
        self.block_n_override = 0
        self._stage: Dict[str, torch.Tensor] = {}
        self._aliased = set()
        self._jobs, self._job_keep, self._jobs_uploaded = [], [], None
        # inference: GroupNorm + SiLU applied inside the consuming conv (dsg_conv gn_coef) wherever the shape allows
        # GroupNorm + SiLU applied inside the consuming conv (dsg_conv gn_coef).  Measured on B200 (profiles/README.md): a
        # LOSS — the conv mainloop already runs at the shared-memory bandwidth limit and the in-place transform of every
        # box (three times per element: one box per column shift) adds smem traffic and one MUFU per element, so the
        # convs slow down by more than the removed GroupNorm pass costs.  Kept as an opt-in for experiments:
        # DSG_FUSE_GN = 0 (default) never, 1 = convs with 256-wide tiles only, 2 = wherever the kernel supports it.
        self.fuse_gn = int(os.environ.get("DSG_FUSE_GN", "0"))
        # samples per L2 group for the full-resolution ops of the sampling program (0 = whole batch per launch):
        # the level-0 tensors of a 16-sample batch are 134 MB each (> the 126 MB L2), those of a 2-4 sample group are
        # not, so running conv -> GroupNorm -> conv ... group by group lets each consumer find its input in L2
        self.l2_group = int(os.environ.get("DSG_L2_GROUP", "0"))
        self.conv_out_mma = int(os.environ.get("DSG_CONV_OUT_MMA", "1"))   # fused conv_norm_out + SiLU + conv_out
        self.train_packs = False  # also pack the data-gradient forms of every conv weight (training path)
        self.train_programs: Dict[Tuple[int, int, int, int], object] = {}
        self._build_topology()

    # ------------------------------------------------------------------ topology (mirrors upstream __init__)
    def _build_topology(self):
        cfg = self.cfg
        boc = list(cfg["block_out_channels"])
        lpb = int(cfg.get("layers_per_block", 2))
        hd = cfg.get("attention_head_dim", 8)
        self.temb_hidden = boc[0] * 4
        self.time_dim = boc[0]
        resnets: List[dict] = []   # every ResnetBlock2D in execution order, with its state-dict prefix
        self.down = []
        out_ch = boc[0]
        for i, t in enumerate(cfg["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            attn = t == "AttnDownBlock2D"
            if t not in ("DownBlock2D", "AttnDownBlock2D"):
                raise ValueError(f"{t} does not exist.")
            blk = {"resnets": [], "attn": [], "down": i != len(boc) - 1, "ch": out_ch, "prefix": f"down_blocks.{i}"}
            for j in range(lpb):
                r = {"prefix": f"down_blocks.{i}.resnets.{j}", "cin": in_ch if j == 0 else out_ch, "cskip": 0,
                     "cout": out_ch}
                blk["resnets"].append(r)
                resnets.append(r)
                if attn:
                    blk["attn"].append({"prefix": f"down_blocks.{i}.attentions.{j}", "ch": out_ch,
                                        "head_dim": hd if hd is not None else out_ch})
            self.down.append(blk)
        mid_ch = boc[-1]
        self.mid = {"resnets": [], "attn": None}
        for j in range(2):
            r = {"prefix": f"mid_block.resnets.{j}", "cin": mid_ch, "cskip": 0, "cout": mid_ch}
            self.mid["resnets"].append(r)
            resnets.append(r)
        if cfg.get("add_attention", True):
            self.mid["attn"] = {"prefix": "mid_block.attentions.0", "ch": mid_ch,
                                "head_dim": hd if hd is not None else mid_ch}
        self.up = []
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, t in enumerate(cfg["up_block_types"]):
            if t not in ("UpBlock2D", "AttnUpBlock2D"):
                raise ValueError(f"{t} does not exist.")
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            attn = t == "AttnUpBlock2D"
            blk = {"resnets": [], "attn": [], "up": i != len(boc) - 1, "ch": out_ch, "prefix": f"up_blocks.{i}"}
            for j in range(lpb + 1):
                skip = in_ch if j == lpb else out_ch
                r_in = prev if j == 0 else out_ch
                r = {"prefix": f"up_blocks.{i}.resnets.{j}", "cin": r_in, "cskip": skip, "cout": out_ch}
                blk["resnets"].append(r)
                resnets.append(r)
                if attn:
                    blk["attn"].append({"prefix": f"up_blocks.{i}.attentions.{j}", "ch": out_ch,
                                        "head_dim": hd if hd is not None else out_ch})
            self.up.append(blk)
        off = 0
        for r in resnets:
            r["temb_off"] = off
            off += r["cout"]
        self.resnets = resnets
        self.proj_total = off
        self.n_levels = len(boc)

    # ------------------------------------------------------------------ weights
    def _src_f32(self, name: str, t: torch.Tensor) -> torch.Tensor:
        """An fp32, contiguous, on-device source for a pack job with a STABLE address: the tensor itself when it
        already is one (parameters of a model living on this device), else a persistent staging copy."""
        t = t.detach()
        if t.device == self.device and t.dtype == torch.float32 and t.is_contiguous():
            return t
        buf = self._stage.get(name)
        if buf is None or buf.shape != t.shape:
            buf = torch.empty(t.shape, dtype=torch.float32, device=self.device)
            self._stage[name] = buf
        buf.copy_(t)
        return buf

    def _pack_conv(self, name: str, mode: int, w: torch.Tensor, w_sc: Optional[torch.Tensor] = None):
        """queue one weight-packing job (all jobs of a load_state_dict run as ONE launch, _run_pack_jobs)."""
        w = self._src_f32(name + "/w", w)
        cout, cin = int(w.shape[0]), int(w.shape[1])
        csc = 0
        if w_sc is not None:
            w_sc = self._src_f32(name + "/wsc", w_sc.detach().reshape(cout, -1))
            csc = int(w_sc.shape[1])
        k = self.lib.dsg_packed_k(mode, cin, csc)
        rows = self.lib.dsg_packed_rows(mode, cout)
        out = self._weight_buf(name, rows * k, torch.float16)
        self._jobs.append((mode, cout, cin, csc, w.data_ptr(), 0 if w_sc is None else w_sc.data_ptr(), out.data_ptr(),
                           int(k), int(rows)))
        self._job_keep.extend([w, w_sc])

    def _pack_dgrad(self, name: str, fwd_mode: int, w: torch.Tensor):
        """data-gradient packing of a conv weight (pack modes 10-13, include/dsg_b200.h)."""
        w = self._src_f32(name + "/w", w)
        cout, cin = int(w.shape[0]), int(w.shape[1])
        k = self.lib.dsg_packed_k_dgrad(fwd_mode, cout)
        rows = self.lib.dsg_packed_rows_dgrad(fwd_mode, cin)
        out = self._weight_buf(name, rows * k, torch.float16)
        self._jobs.append((10 + fwd_mode, cout, cin, 0, w.data_ptr(), 0, out.data_ptr(), int(k), int(rows)))
        self._job_keep.append(w)

    def _run_pack_jobs(self):
        """ONE kernel launch packs every conv weight (dsg_pack_conv_weights_batched); the device-side job table is
        rebuilt only when a pointer or shape changed."""
        if not self._jobs:
            return
        if self._jobs != self._jobs_uploaded:
            batched, single, chunk = [], [], 0
            for job in self._jobs:
                mode, cout, cin, csc = job[:4]
                nb = int(self.lib.dsg_pack_job_blocks(mode, cout, cin, csc))
                if nb < 0:          # a source row too wide for the shared-memory staged form: element-wise kernel
                    single.append(job)
                else:
                    batched.append((job, chunk))
                    chunk += nb
            arr = (_lib.PackJob * max(len(batched), 1))()
            for j, ((mode, cout, cin, csc, w, wsc, out, k, rows), begin) in zip(arr, batched):
                j.mode, j.cout, j.cin, j.csc = mode, cout, cin, csc
                j.w, j.w_sc, j.out = w, (wsc or None), out
                j.k_total, j.rows, j.chunk_begin = k, rows, begin
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._jobs_dev = raw.to(self.device)
            self._jobs_blocks, self._jobs_batched, self._jobs_single = chunk, len(batched), single
            self._jobs_uploaded = list(self._jobs)
        st = torch.cuda.current_stream(self.device).cuda_stream
        if self._jobs_batched:
            check(self.lib.dsg_pack_conv_weights_batched(self._jobs_dev.data_ptr(), self._jobs_batched,
                                                         self._jobs_blocks, st), "pack weights (batched)")
        for mode, cout, cin, csc, w, wsc, out, k, rows in self._jobs_single:
            check(self.lib.dsg_pack_conv_weight(mode, w, cout, cin, wsc or None, csc, out, st), "pack weight")

    def _weight_buf(self, name: str, numel: int, dtype) -> torch.Tensor:
        """Packed-weight buffers keep their address across re-packs (a training step re-packs every step, and the
        execution programs hold raw pointers); a new or resized buffer invalidates the programs."""
        t = self.weights.get(name)
        if t is None or t.numel() != numel or t.dtype != dtype or name in self._aliased:
            t = torch.empty(numel, dtype=dtype, device=self.device)
            self.weights[name] = t
            self._aliased.discard(name)
            self._new_weight_buffers = True
        return t

    def _f32(self, name: str, t: torch.Tensor, alias: bool = True):
        """fp32 side tensors (GroupNorm affines, biases, ...).  A parameter that already lives on this device as a
        contiguous fp32 tensor is used IN PLACE (no copy; optimizers update it in place); anything else is copied into a
        persistent buffer."""
        t = t.detach()
        if alias and t.device == self.device and t.dtype == torch.float32 and t.is_contiguous() and t.numel() > 0 \
                and t.data_ptr() % 16 == 0:
            old = self.weights.get(name)
            if old is None or old.data_ptr() != t.data_ptr() or old.shape != t.shape:
                self._new_weight_buffers = True
            self.weights[name] = t
            self._aliased.add(name)
            return
        buf = self._weight_buf(name, t.numel(), torch.float32)
        if buf.shape != t.shape:
            buf = buf.view(t.shape)
            self.weights[name] = buf
        buf.copy_(t)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """(Re)pack all weights from a state dict with upstream key names."""
        self._new_weight_buffers = False
        self._jobs, self._job_keep = [], []
        with torch.cuda.device(self.device):
            self._f32("conv_in.w", sd["conv_in.weight"])
            self._f32("conv_in.b", sd["conv_in.bias"])
            self._f32("conv_out.w", sd["conv_out.weight"])
            self._f32("conv_out.b", sd["conv_out.bias"])
            # tensor-core form of conv_out: output channels zero-padded to the smallest UMMA N (16)
            wo = sd["conv_out.weight"].detach()
            if wo.shape[0] <= 16 and wo.shape[1] % 64 == 0:
                w16 = self._stage.get("conv_out.w16")
                if w16 is None or w16.shape[1:] != wo.shape[1:]:
                    w16 = torch.zeros((16,) + tuple(wo.shape[1:]), dtype=torch.float32, device=self.device)
                    self._stage["conv_out.w16"] = w16
                    self._stage["conv_out.b16"] = torch.zeros(16, dtype=torch.float32, device=self.device)
                b16 = self._stage["conv_out.b16"]
                w16[: wo.shape[0]].copy_(wo)
                b16[: wo.shape[0]].copy_(sd["conv_out.bias"].detach())
                self._pack_conv("conv_out.w16", 0, w16)
                self.weights["conv_out.b16"] = b16
            self._f32("norm_out.g", sd["conv_norm_out.weight"])
            self._f32("norm_out.b", sd["conv_norm_out.bias"])
            for k in ("linear_1", "linear_2"):
                self._f32(f"te.{k}.w", sd[f"time_embedding.{k}.weight"].t())   # [in][out]: see dsg_time_embed
                self._f32(f"te.{k}.b", sd[f"time_embedding.{k}.bias"])
            for nm, key in (("te.proj.w", ".time_emb_proj.weight"), ("te.proj.b", ".time_emb_proj.bias")):
                parts = [sd[r["prefix"] + key].detach().to(self.device, torch.float32) for r in self.resnets]
                shape = (sum(p.shape[0] for p in parts),) + tuple(parts[0].shape[1:])
                buf = self._stage.get(nm)
                if buf is None or buf.shape != shape:
                    buf = torch.empty(shape, dtype=torch.float32, device=self.device)
                    self._stage[nm] = buf
                torch.cat(parts, 0, out=buf)
                self._f32(nm, buf)
            half = self.time_dim // 2
            # exp table computed on the host exactly like upstream get_timestep_embedding (models/embeddings.py).  It
            # depends on the configuration only: uploaded once — a blocking host-to-device copy per re-pack would make
            # the host wait for the whole backward pass of a training step and leave the GPU idle while the rest of this
            # function runs (1.3 ms per step, measured)
            fkey = (half, float(self.cfg.get("freq_shift", 0)))
            if getattr(self, "_freqs_key", None) != fkey or "te.freqs" not in self.weights:
                exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
                exponent = exponent / (half - fkey[1])
                self._f32("te.freqs", torch.exp(exponent))
                self._freqs_key = fkey
            for r in self.resnets:
                pre = r["prefix"]
                for nm in ("norm1", "norm2"):
                    self._f32(f"{pre}.{nm}.g", sd[f"{pre}.{nm}.weight"])
                    self._f32(f"{pre}.{nm}.b", sd[f"{pre}.{nm}.bias"])
                self._pack_conv(f"{pre}.conv1", 0, sd[f"{pre}.conv1.weight"])
                self._f32(f"{pre}.conv1.b", sd[f"{pre}.conv1.bias"])
                sc_key = f"{pre}.conv_shortcut.weight"
                has_sc = sc_key in sd
                if has_sc != (r["cin"] + r["cskip"] != r["cout"]):
                    raise DsgError(f"{pre}: conv_shortcut presence does not match channel counts")
                # cin == cout blocks have a plain residual.  For narrow outputs (cout <= 128) it rides through the
                # tensor core as an identity 1x1 shortcut panel (x * 1.0 is exact in fp16, the sum is fp32): the
                # residual tile then arrives by TMA like every other operand instead of through latency-bound
                # global loads in the epilogue, which those short-K tiles cannot hide.
                r["id_sc"] = (not has_sc) and r["cout"] <= 128
                if r["id_sc"]:
                    if not hasattr(self, "_eye"):
                        self._eye = {}
                    if r["cout"] not in self._eye:
                        self._eye[r["cout"]] = torch.eye(r["cout"], dtype=torch.float32, device=self.device)
                    w_sc = self._eye[r["cout"]]
                else:
                    w_sc = sd[sc_key] if has_sc else None
                self._pack_conv(f"{pre}.conv2", 0, sd[f"{pre}.conv2.weight"], w_sc)
                b2 = sd[f"{pre}.conv2.bias"].detach().to(self.device, torch.float32)
                if has_sc:
                    b2 = b2 + sd[f"{pre}.conv_shortcut.bias"].detach().to(self.device, torch.float32)
                self._f32(f"{pre}.conv2.b", b2, alias=not has_sc)   # with a shortcut it is a derived sum: copy
                r["has_sc"] = has_sc
                if self.train_packs:
                    self._pack_dgrad(f"{pre}.conv1.dg", 0, sd[f"{pre}.conv1.weight"])
                    self._pack_dgrad(f"{pre}.conv2.dg", 0, sd[f"{pre}.conv2.weight"])
                    if has_sc:
                        self._pack_dgrad(f"{pre}.sc.dg", 3, sd[sc_key])
            attns = [a for blk in self.down for a in blk["attn"]] + ([self.mid["attn"]] if self.mid["attn"] else []) \
                + [a for blk in self.up for a in blk["attn"]]
            for a in attns:
                pre = a["prefix"]
                self._f32(f"{pre}.gn.g", sd[f"{pre}.group_norm.weight"])
                self._f32(f"{pre}.gn.b", sd[f"{pre}.group_norm.bias"])
                parts = [sd[f"{pre}.to_{n}.weight"].detach().to(self.device, torch.float32) for n in "qkv"]
                wqkv = self._stage.get(f"{pre}.wqkv")
                if wqkv is None:
                    wqkv = torch.empty((3 * parts[0].shape[0], parts[0].shape[1], 1, 1), dtype=torch.float32,
                                       device=self.device)
                    self._stage[f"{pre}.wqkv"] = wqkv
                torch.cat(parts, 0, out=wqkv.view(wqkv.shape[0], wqkv.shape[1]))
                bqkv = torch.cat([sd[f"{pre}.to_q.bias"], sd[f"{pre}.to_k.bias"], sd[f"{pre}.to_v.bias"]], 0)
                self._pack_conv(f"{pre}.qkv", 3, wqkv)
                self._f32(f"{pre}.qkv.b", bqkv, alias=False)
                self._pack_conv(f"{pre}.out", 3, sd[f"{pre}.to_out.0.weight"].detach()[:, :, None, None])
                self._f32(f"{pre}.out.b", sd[f"{pre}.to_out.0.bias"])
                if self.train_packs:
                    self._pack_dgrad(f"{pre}.qkv.dg", 3, wqkv)
                    self._pack_dgrad(f"{pre}.out.dg", 3, sd[f"{pre}.to_out.0.weight"].detach()[:, :, None, None])
            for i, blk in enumerate(self.down):
                if blk["down"]:
                    pre = f"down_blocks.{i}.downsamplers.0.conv"
                    self._pack_conv(pre, 1, sd[pre + ".weight"])
                    self._f32(pre + ".b", sd[pre + ".bias"])
                    if self.train_packs:
                        self._pack_dgrad(pre + ".dg", 1, sd[pre + ".weight"])
            for i, blk in enumerate(self.up):
                if blk["up"]:
                    pre = f"up_blocks.{i}.upsamplers.0.conv"
                    self._pack_conv(pre, 2, sd[pre + ".weight"])
                    self._f32(pre + ".b", sd[pre + ".bias"])
                    if self.train_packs:
                        self._pack_dgrad(pre + ".dg", 2, sd[pre + ".weight"])
            if self.train_packs:
                for k in ("linear_1", "linear_2"):
                    self._f32(f"te.{k}.w_oi", sd[f"time_embedding.{k}.weight"])   # [out][in] for the backward
            self._run_pack_jobs()
        self.packed = True
        if self._new_weight_buffers:   # programs hold raw weight pointers
            self.programs.clear()
            self.train_programs.clear()

    # ------------------------------------------------------------------ program construction
    def program(self, batch: int, h: int, w: int) -> "_Program":
        key = (batch, h, w)
        prog = self.programs.get(key)
        if prog is None:
            if not self.packed:
                raise DsgError("UNetEngine: weights have not been loaded")
            div = 2 ** (self.n_levels - 1)
            if h % div or w % div:
                raise DsgError(f"input {h}x{w} must be divisible by {div}")
            prog = _Program(self, batch, h, w)
            self.programs[key] = prog
        return prog

    def train_program(self, batch: int, h: int, w: int, grad_slices: Dict[str, torch.Tensor], slot: int = 0):
        """forward-with-saved-activations + backward program (engine_train.TrainProgram) writing parameter
        gradients into grad_slices (name -> fp32 view); one program per (shape, gradient-buffer slot)."""
        from .engine_train import TrainProgram
        key = (batch, h, w, slot)
        prog = self.train_programs.get(key)
        if prog is None:
            if not self.packed or not self.train_packs:
                raise DsgError("UNetEngine: training weights have not been packed")
            div = 2 ** (self.n_levels - 1)
            if h % div or w % div:
                raise DsgError(f"input {h}x{w} must be divisible by {div}")
            prog = TrainProgram(self, batch, h, w, grad_slices)
            self.train_programs[key] = prog
        return prog

    def forward(self, sample: torch.Tensor, t_float: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """sample: fp32 NCHW on this device; t_float: fp32 [B] timestep values on this device."""
        b, c, h, w = sample.shape
        prog = self.program(b, h, w)
        return prog.run(sample, t_float, out)


class _Program:
    """The flat kernel sequence of one forward pass for a fixed (batch, H, W)."""

    def __init__(self, eng: UNetEngine, batch: int, h: int, w: int):
        self.eng, self.b, self.h, self.w = eng, batch, h, w
        self.lib = eng.lib
        self.ops: List[Callable[[int], None]] = []
        self.op_info: List[Tuple[str, dict]] = []
        self.keep: list = []          # ctypes structs / tensors that must outlive the closures
        self.uid = 0
        self.in_ptr = C.c_void_p(0)   # set per run (or to a static buffer under graph capture)
        self.t_ptr = C.c_void_p(0)
        self.out_ptr = C.c_void_p(0)
        self.cin = int(eng.cfg.get("in_channels", 3))
        self.cout = int(eng.cfg.get("out_channels", 3))
        self.fuse_gn = getattr(self, "fuse_gn", True)   # the training program keeps the normalised tensors: no fusion
        # pass 1 sizes the shared temporaries (so no buffer is outgrown mid-program), pass 2 emits the ops
        self._measure: Optional[Dict[Tuple[str, torch.dtype], int]] = {}
        self._stats_off = 0
        self.stats_all: Optional[torch.Tensor] = None   # every GroupNorm-statistics slice of this program, contiguous
        self.stats_of: Dict[int, torch.Tensor] = {}     # activation buffer (data_ptr) -> its int64 [b][C][2] totals
        self._build()
        for (name, dtype), numel in self._measure.items():
            eng.arena.get(name, numel, dtype)
        self._measure = None
        self.stats_all = eng.arena.get(f"{batch}x{h}x{w}/gn_stats", max(self._stats_off, 2), torch.int64)
        self._stats_off = 0
        self.stats_of = {}
        self.ops, self.keep, self.op_info, self.op_sub = [], [], [], []
        self._build()
        g = int(getattr(eng, "l2_group", 0))
        if self.regroup and g and g < self.b and self.b % g == 0:
            self._regroup(g)

    regroup = True   # the training program (engine_train.TrainProgram) keeps whole-batch launches
    fuse_out_mma = True   # ... and the activated conv_norm_out tensor (conv_out's weight gradient reads it)

    def _regroup(self, g: int):
        """Re-issue every run of consecutive full-resolution ops sample-group by sample-group (depth first): all of
        the U-Net is independent per sample, so the result is bit-identical; only the order of launches changes."""
        full = self.h * self.w
        ops, info = [], []
        i, n = 0, len(self.ops)
        while i < n:
            j = i
            while j < n and self.op_sub[j][0] is not None and self.op_sub[j][1] >= full:
                j += 1
            if j - i < 2:
                ops.append(self.ops[i]); info.append(self.op_info[i])
                i += 1
                continue
            for n0 in range(0, self.b, g):
                for k in range(i, j):
                    name, meta = self.op_info[k]
                    ops.append(self.op_sub[k][0](n0, g))
                    info.append((name, {kk: (v * g // self.b if kk in ("flops", "flops_exec", "bytes") else v)
                                        for kk, v in meta.items()}))
            i = j
        self.ops, self.op_info = ops, info
        self.n_launches = len(self.ops) + 1

    # --- buffers -------------------------------------------------------------------------------------
    def _stats(self, ch: int) -> torch.Tensor:
        """A fresh int64 [b][ch][2] slice of per-channel GroupNorm totals (zeroed once per run with all the others)."""
        n = self.b * ch * 2
        off = self._stats_off
        self._stats_off += n
        if self.stats_all is None:     # measuring pass
            return self.eng.arena.get("tmp/_measure_i64", 16, torch.int64)
        return self.stats_all[off:off + n]

    def _new(self, name: str, hw: Tuple[int, int], ch: int) -> torch.Tensor:
        """A persistent activation buffer; whichever conv writes it also accumulates its GroupNorm statistics."""
        numel = self.b * hw[0] * hw[1] * ch
        t = self.eng.arena.get(f"{self.b}x{self.h}x{self.w}/{name}", numel, torch.float16)
        self.stats_of[t.data_ptr()] = self._stats(ch)
        return t

    def _tmp_raw(self, name: str, numel: int, dtype) -> torch.Tensor:
        # temporaries are shared by name across layers (and shapes): sized to the largest request
        if self._measure is not None:
            key = (f"tmp/{name}", dtype)
            self._measure[key] = max(self._measure.get(key, 0), numel)
            return self.eng.arena.get("tmp/_measure", 16, dtype)
        return self.eng.arena.get(f"tmp/{name}", numel, dtype)

    def _tmp(self, name: str, hw: Tuple[int, int], ch: int) -> torch.Tensor:
        return self._tmp_raw(name, self.b * hw[0] * hw[1] * ch, torch.float16)

    def _te_saved(self, half: int, hidden: int) -> Optional[torch.Tensor]:
        return None   # inference keeps no time-embedding activations (the training program does)

    # --- op emitters ---------------------------------------------------------------------------------
    def _emit(self, name: str, meta: dict, fn: Callable[[int], None], sub=None, px: int = 0):
        """sub(n0, ng) -> the same op restricted to samples [n0, n0 + ng) (or None); px = pixels per sample it touches."""
        self.ops.append(fn)
        self.op_info.append((name, meta))
        if hasattr(self, "op_sub"):
            self.op_sub.append((sub, px))

    def _gn_stats_for(self, x1, c1, x2, c2, hw, st1=None, st2=None):
        """the statistics a GroupNorm over cat(x1, x2) reads: accumulated by whoever produced x1 / x2, or — odd group
        sizes, where the conv epilogues' channel-pair totals do not line up with the groups — one extra read."""
        eng, lib, b = self.eng, self.lib, self.b
        npx = hw[0] * hw[1]
        st1 = st1 if st1 is not None else self.stats_of[x1.data_ptr()]
        if x2 is not None and st2 is None:
            st2 = self.stats_of[x2.data_ptr()]
        if ((c1 + c2) // eng.groups) % 2:
            srcs = [(x1, c1)] + ([(x2, c2)] if x2 is not None else [])
            fresh = []
            for xs, cs in srcs:
                stx = self._stats(cs)
                fresh.append(stx)
                ga = (xs.data_ptr(), cs, stx.data_ptr(), b, npx)
                self._emit("gn_stats", {"bytes": b * npx * cs * 2},
                           lambda st, a=ga: check(lib.dsg_gn_stats(*a, st), "gn_stats"))
            st1, st2 = fresh[0], (fresh[1] if len(fresh) > 1 else None)
        return st1, st2

    def _gn(self, x1, c1, x2, c2, hw, gname, bname, act, out, st1=None, st2=None):
        """GroupNorm(+SiLU) of cat(x1, x2); the statistics were accumulated by whoever produced x1 / x2."""
        eng, lib, b = self.eng, self.lib, self.b
        npx = hw[0] * hw[1]
        st1, st2 = self._gn_stats_for(x1, c1, x2, c2, hw, st1, st2)
        g, bt = eng.weights[gname], eng.weights[bname]
        a2 = (_p(x1), c1, st1.data_ptr(), _p(x2), c2, _p(st2), g.data_ptr(), bt.data_ptr(), eng.eps, act,
              out.data_ptr(), b, npx, eng.groups)
        nbytes = b * npx * (c1 + c2) * 2

        def sub(n0, ng, a=a2):
            a = list(a)
            a[0] += n0 * npx * c1 * 2
            a[2] += n0 * c1 * 16
            if x2 is not None:
                a[3] += n0 * npx * c2 * 2
                a[5] += n0 * c2 * 16
            a[10] += n0 * npx * (c1 + c2) * 2
            a[11] = ng
            a = tuple(a)
            return lambda st: check(lib.dsg_gn_apply(*a, st), "gn_apply")
        self._emit("gn_apply", {"bytes": 2 * nbytes}, lambda st, a=a2: check(lib.dsg_gn_apply(*a, st), "gn_apply"),
                   sub=sub, px=npx)
        return st1, st2

    def _gn_coef(self, x1, c1, x2, c2, hw, gname, bname, st1=None, st2=None) -> torch.Tensor:
        """fused form: only the per-(sample, channel) coefficients are computed here; the consuming conv applies
        GroupNorm + SiLU to its activation boxes in shared memory (dsg_conv gn_coef)."""
        eng, lib, b = self.eng, self.lib, self.b
        npx = hw[0] * hw[1]
        st1, st2 = self._gn_stats_for(x1, c1, x2, c2, hw, st1, st2)
        coef = self._tmp_raw("gn_coef", b * (c1 + c2) * 2, torch.float32)
        g, bt = eng.weights[gname], eng.weights[bname]
        a2 = (c1, st1.data_ptr(), c2, _p(st2), g.data_ptr(), bt.data_ptr(), eng.eps, coef.data_ptr(), b, npx,
              eng.groups)
        self._emit("gn_coef", {"bytes": b * (c1 + c2) * 24}, lambda st, a=a2: check(lib.dsg_gn_coef(*a, st), "gn_coef"))
        return coef

    def _fusable(self, hw, cin, cin1, cout) -> bool:
        eng = self.eng
        if not (self.fuse_gn and eng.fuse_gn and eng.conv_impl != 1):
            return False
        if eng.fuse_gn == 1 and cout % 256 != 0:
            return False
        return bool(self.lib.dsg_conv_gn_fusable(0, hw[0], hw[1], cin, cin1, cout))

    def _conv(self, mode, x, hw, cin, cout, wname, bname, out, temb_off=None, residual=None, sc1=None, csc1=0,
              sc2=None, csc2=0, count_sc=True, stats=None, x2=None, cin1=0, gn_coef=None):
        eng, lib = self.eng, self.lib
        a = ConvArgs()
        a.mode, a.n, a.h, a.w, a.cin, a.cout = mode, self.b, hw[0], hw[1], cin, cout
        a.x = x.data_ptr()
        a.sc1, a.csc1, a.sc2, a.csc2 = _p(sc1), csc1, _p(sc2), csc2
        a.wpacked = eng.weights[wname].data_ptr()
        a.bias = eng.weights[bname].data_ptr()
        if temb_off is not None:
            a.temb, a.temb_stride, a.temb_off = self.temb.data_ptr(), eng.proj_total, temb_off
        a.residual = _p(residual)
        a.out = out.data_ptr()
        stats = stats if stats is not None else self.stats_of.get(out.data_ptr())
        if stats is not None:
            a.out_stats = stats.data_ptr()   # the epilogue accumulates the next GroupNorm's statistics
        a.block_n = eng.block_n_override if (eng.block_n_override and cout % eng.block_n_override == 0) else 0
        a.impl = eng.conv_impl
        if gn_coef is not None:   # fused GroupNorm + SiLU (+ concat) on the input: x / x2 are the RAW tensors
            a.x2, a.cin1, a.gn_coef = _p(x2), (cin1 or cin), gn_coef.data_ptr()
        self.keep.append(a)
        ref = C.byref(a)
        # reference op count: an identity shortcut panel stands for a residual ADD, not for GEMM work
        k_ref = {0: 9 * cin + (csc1 + csc2 if count_sc else 0), 1: 9 * cin, 2: 9 * cin, 3: cin}[mode]
        opx = {0: hw[0] * hw[1], 1: hw[0] * hw[1] // 4, 2: hw[0] * hw[1] * 4, 3: hw[0] * hw[1]}[mode]
        # executed: the upsample conv runs as 4 sub-pixel phases of 2x2 taps (4/9 of the MACs), an identity shortcut
        # panel is real GEMM work
        k_exec = {0: 9 * cin + csc1 + csc2, 1: 9 * cin, 2: 4 * cin, 3: cin}[mode]
        meta = {"mode": mode, "hw": hw, "cin": cin + csc1 + csc2, "cout": cout,
                "flops": 2 * self.b * opx * cout * k_ref,   # algorithmic (reference op count, no sub-pixel discount)
                "flops_exec": 2 * self.b * opx * cout * k_exec}
        sub = None
        if gn_coef is None:
            def sub(n0, ng, a=a):
                return self._conv_sub(a, n0, ng, opx, f"conv {wname}")
        self._emit("conv", meta, lambda st, r=ref: check(lib.dsg_conv(r, st), f"conv {wname}"), sub=sub,
                   px=max(opx, hw[0] * hw[1]))

    def _conv_sub(self, a: ConvArgs, n0: int, ng: int, opx: int, what: str, out_f32_of=None):
        """dsg_conv of `a` restricted to samples [n0, n0 + ng): a copy of the argument block with shifted pointers."""
        lib = self.lib
        s = ConvArgs.from_buffer_copy(a)
        ipx = a.h * a.w
        s.n = ng
        s.x = a.x + n0 * ipx * a.cin * 2
        if a.sc1:
            s.sc1 = a.sc1 + n0 * ipx * a.csc1 * 2
        if a.sc2:
            s.sc2 = a.sc2 + n0 * ipx * a.csc2 * 2
        if a.temb:
            s.temb = a.temb + n0 * a.temb_stride * 4
        if a.residual:
            s.residual = a.residual + n0 * opx * a.cout * 2
        if a.out:
            s.out = a.out + n0 * opx * a.cout * 2
        if a.out_stats:
            s.out_stats = a.out_stats + n0 * a.cout * 16
        self.keep.append(s)
        r = C.byref(s)
        if out_f32_of is None:
            return lambda st: check(lib.dsg_conv(r, st), what)
        off = n0 * a.cout_real * opx * 4

        def run(st):
            s.out_nchw_f32 = out_f32_of().value + off
            check(lib.dsg_conv(r, st), what)
        return run

    def _resnet(self, r, x1, x2, hw, out):
        c1, c2, co, pre = r["cin"], r["cskip"], r["cout"], r["prefix"]
        hbuf = self._tmp("h", hw, co)
        h_stats = self._stats(co)   # the "h" buffer is shared between blocks, its statistics are not
        act = act2 = None
        if self._fusable(hw, c1 + c2, c1, co):
            coef = self._gn_coef(x1, c1, x2, c2, hw, f"{pre}.norm1.g", f"{pre}.norm1.b")
            st1 = st2 = None
            self._conv(0, x1, hw, c1 + c2, co, f"{pre}.conv1", f"{pre}.conv1.b", hbuf, temb_off=r["temb_off"],
                       stats=h_stats, x2=x2, cin1=c1, gn_coef=coef)
        else:
            act = self._tmp("act", hw, c1 + c2)
            st1, st2 = self._gn(x1, c1, x2, c2, hw, f"{pre}.norm1.g", f"{pre}.norm1.b", 1, act)
            self._conv(0, act, hw, c1 + c2, co, f"{pre}.conv1", f"{pre}.conv1.b", hbuf, temb_off=r["temb_off"],
                       stats=h_stats)
        kw = {}
        if self._fusable(hw, co, co, co):
            kw = dict(gn_coef=self._gn_coef(hbuf, co, None, 0, hw, f"{pre}.norm2.g", f"{pre}.norm2.b", st1=h_stats))
            src = hbuf
        else:
            act2 = self._tmp("act", hw, co)
            self._gn(hbuf, co, None, 0, hw, f"{pre}.norm2.g", f"{pre}.norm2.b", 1, act2, st1=h_stats)
            src = act2
        if r["has_sc"] or r["id_sc"]:
            self._conv(0, src, hw, co, co, f"{pre}.conv2", f"{pre}.conv2.b", out, sc1=x1, csc1=c1, sc2=x2, csc2=c2,
                       count_sc=r["has_sc"], **kw)
        else:
            self._conv(0, src, hw, co, co, f"{pre}.conv2", f"{pre}.conv2.b", out, residual=x1, **kw)
        return {"kind": "resnet", "r": r, "x1": x1, "x2": x2, "hw": hw, "out": out, "a1": act, "h": hbuf,
                "h_stats": h_stats, "a2": act2, "st1": st1, "st2": st2}

    def _attn(self, a, x, hw, out):
        eng, lib, b = self.eng, self.lib, self.b
        ch, pre, hd = a["ch"], a["prefix"], a["head_dim"]
        act = self._tmp("act", hw, ch)
        st1, _ = self._gn(x, ch, None, 0, hw, f"{pre}.gn.g", f"{pre}.gn.b", 0, act)
        qkv = self._tmp("qkv", hw, 3 * ch)
        self._conv(3, act, hw, ch, 3 * ch, f"{pre}.qkv", f"{pre}.qkv.b", qkv)
        o = self._tmp("attn_o", hw, ch)
        ntok = hw[0] * hw[1]
        lse = self._attn_lse(b, ch // hd, ntok, hd)
        meta = {"flops": 4 * b * ntok * ntok * ch, "exps": b * (ch // hd) * ntok * ntok}
        if lse is not None:   # training forward: the tcgen05 kernel also leaves the log-sum-exp for the backward
            args = (qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), b, ntok, ch // hd, hd)
            self._emit("attention", meta, lambda st, a_=args: check(lib.dsg_attention_train(*a_, st), "attention"))
        else:
            args = (qkv.data_ptr(), o.data_ptr(), b, ntok, ch // hd, hd)
            self._emit("attention", meta, lambda st, a_=args: check(lib.dsg_attention(*a_, st), "attention"))
        self._conv(3, o, hw, ch, ch, f"{pre}.out", f"{pre}.out.b", out, residual=x)
        return {"kind": "attn", "a": a, "x": x, "hw": hw, "out": out, "act": act, "qkv": qkv, "o": o, "st1": st1,
                "lse": lse}

    def _attn_lse(self, b: int, heads: int, tokens: int, hd: int) -> Optional[torch.Tensor]:
        return None   # inference keeps no softmax statistics (the training program does)

    def _record(self, rec: dict):
        """hook: the training program keeps what each block produced (engine_train.py); inference drops it."""

    # --- whole forward -------------------------------------------------------------------------------
    def _build(self):
        eng, lib, b = self.eng, self.lib, self.b
        W = eng.weights
        hw = (self.h, self.w)
        self.temb = eng.arena.get(f"{b}/temb", b * eng.proj_total, torch.float32)
        emb_ws = eng.arena.get(f"{b}/emb_ws", b * eng.temb_hidden, torch.float32)
        half = eng.time_dim // 2
        flip = 1 if eng.cfg.get("flip_sin_to_cos", True) else 0
        te_args = (W["te.freqs"].data_ptr(), half, flip, W["te.linear_1.w"].data_ptr(), W["te.linear_1.b"].data_ptr(),
                   W["te.linear_2.w"].data_ptr(), W["te.linear_2.b"].data_ptr(), eng.temb_hidden,
                   W["te.proj.w"].data_ptr(), W["te.proj.b"].data_ptr(), eng.proj_total, emb_ws.data_ptr(),
                   self.temb.data_ptr(), b)
        self._emit("zero_stats", {}, lambda st: self.stats_all.zero_())
        te_saved = self._te_saved(half, eng.temb_hidden)
        self.emb_ws = emb_ws
        self._emit("time_embed", {}, lambda st: check(lib.dsg_time_embed_ex(self.t_ptr, *te_args, _p(te_saved), st),
                                                      "time_embed"))
        c0 = eng.cfg["block_out_channels"][0]
        x = self._new("conv_in", hw, c0)
        # conv_in writes the per-channel GroupNorm totals of its output itself (tensor-core form; other widths fall
        # back to the CUDA-core kernel + one statistics pass inside the same entry point)
        ci_args = (W["conv_in.w"].data_ptr(), W["conv_in.b"].data_ptr(), x.data_ptr(),
                   self.stats_of[x.data_ptr()].data_ptr(), b, self.cin, hw[0], hw[1], c0)
        def conv_in_sub(n0, ng, a=ci_args):
            npx = hw[0] * hw[1]
            a2 = (a[0], a[1], a[2] + n0 * npx * c0 * 2, a[3] + n0 * c0 * 16, ng) + a[5:]
            off = n0 * self.cin * npx * 4
            return lambda st: check(lib.dsg_conv_in_stats(C.c_void_p(self.in_ptr.value + off), *a2, st), "conv_in")
        self._emit("conv_in", {"bytes": b * hw[0] * hw[1] * (self.cin * 4 + c0 * 2)},
                   lambda st: check(lib.dsg_conv_in_stats(self.in_ptr, *ci_args, st), "conv_in"),
                   sub=conv_in_sub, px=hw[0] * hw[1])
        self._record({"kind": "in", "out": x, "hw": hw, "c0": c0})
        skips = [(x, c0, hw)]
        for i, blk in enumerate(eng.down):
            for j, r in enumerate(blk["resnets"]):
                has_attn = bool(blk["attn"])
                out = self._new(f"d{i}r{j}" + ("pre" if has_attn else ""), hw, r["cout"])
                self._record(self._resnet(r, x, None, hw, out))
                x = out
                if has_attn:
                    out = self._new(f"d{i}a{j}", hw, r["cout"])
                    self._record(self._attn(blk["attn"][j], x, hw, out))
                    x = out
                skips.append((x, r["cout"], hw))
            if blk["down"]:
                nhw = (hw[0] // 2, hw[1] // 2)
                out = self._new(f"d{i}ds", nhw, blk["ch"])
                pre = f"down_blocks.{i}.downsamplers.0.conv"
                self._conv(1, x, hw, blk["ch"], blk["ch"], pre, pre + ".b", out)
                self._record({"kind": "down", "prefix": pre, "x": x, "hw": hw, "out": out, "ch": blk["ch"]})
                x, hw = out, nhw
                skips.append((x, blk["ch"], hw))
        mid = eng.mid
        self._record({"kind": "boundary"})   # backward: everything recorded after this point is done when it is reached
        out = self._new("m0", hw, mid["resnets"][0]["cout"])
        self._record(self._resnet(mid["resnets"][0], x, None, hw, out))
        x = out
        if mid["attn"] is not None:
            out = self._new("ma", hw, mid["attn"]["ch"])
            self._record(self._attn(mid["attn"], x, hw, out))
            x = out
        out = self._new("m1", hw, mid["resnets"][1]["cout"])
        self._record(self._resnet(mid["resnets"][1], x, None, hw, out))
        x = out
        for i, blk in enumerate(eng.up):
            for j, r in enumerate(blk["resnets"]):
                sk, sk_c, sk_hw = skips.pop()
                assert sk_c == r["cskip"] and sk_hw == hw, (sk_c, r["cskip"], sk_hw, hw)
                has_attn = bool(blk["attn"])
                out = self._new(f"u{i}r{j}" + ("pre" if has_attn else ""), hw, r["cout"])
                self._record(self._resnet(r, x, sk, hw, out))
                x = out
                if has_attn:
                    out = self._new(f"u{i}a{j}", hw, r["cout"])
                    self._record(self._attn(blk["attn"][j], x, hw, out))
                    x = out
            if blk["up"]:
                nhw = (hw[0] * 2, hw[1] * 2)
                out = self._new(f"u{i}us", nhw, blk["ch"])
                pre = f"up_blocks.{i}.upsamplers.0.conv"
                self._conv(2, x, hw, blk["ch"], blk["ch"], pre, pre + ".b", out)
                self._record({"kind": "up", "prefix": pre, "x": x, "hw": hw, "out": out, "ch": blk["ch"]})
                x, hw = out, nhw
        assert not skips
        tw = 16 if hw[1] >= 16 else (8 if hw[1] >= 8 else 0)
        tc_out = bool("conv_out.w16" in W and tw and hw[0] >= 2 * (128 // tw) + 2 and eng.conv_impl != 1)
        fuse_out = tc_out and self._fusable(hw, c0, c0, 16)
        coef = act = st_out = None
        meta = {"bytes": b * hw[0] * hw[1] * (self.cout * 4 + c0 * 2), "flops": 2 * b * hw[0] * hw[1] * self.cout * 9 * c0}
        if self.fuse_out_mma and eng.conv_out_mma and c0 == 64 and self.cout <= 8 and eng.conv_impl != 1:
            # inference: conv_norm_out + SiLU + conv_out in ONE pass over the raw tensor (warp-level mma.sync kernel)
            coef = self._gn_coef(x, c0, None, 0, hw, "norm_out.g", "norm_out.b")
            self._record({"kind": "out", "x": x, "hw": hw, "act": None, "st1": None, "c0": c0})
            cf_args = (x.data_ptr(), coef.data_ptr(), W["conv_out.w"].data_ptr(), W["conv_out.b"].data_ptr())
            cf_tail = (b, c0, hw[0], hw[1], self.cout)
            self._emit("conv_out", meta,
                       lambda st: check(lib.dsg_conv_out_fused(*cf_args, self.out_ptr, *cf_tail, st), "conv_out (fused)"))
            self.n_launches = len(self.ops) + 1
            return
        if fuse_out:
            coef = self._gn_coef(x, c0, None, 0, hw, "norm_out.g", "norm_out.b")
        else:
            act = self._tmp("act", hw, c0)
            st_out, _ = self._gn(x, c0, None, 0, hw, "norm_out.g", "norm_out.b", 1, act)
        self._record({"kind": "out", "x": x, "hw": hw, "act": act, "st1": st_out, "c0": c0})
        if tc_out:
            # tcgen05 path: the halo-reuse igemm with BLOCK_N = 16 and an NCHW fp32 epilogue
            a = ConvArgs()
            a.mode, a.n, a.h, a.w, a.cin, a.cout = 0, b, hw[0], hw[1], c0, 16
            a.x = (x if fuse_out else act).data_ptr()
            a.wpacked = W["conv_out.w16"].data_ptr()
            a.bias = W["conv_out.b16"].data_ptr()
            a.cout_real = self.cout
            if fuse_out:
                a.cin1, a.gn_coef = c0, coef.data_ptr()
            self.keep.append(a)

            def run_conv_out(st, a=a):
                a.out_nchw_f32 = self.out_ptr.value
                check(lib.dsg_conv(C.byref(a), st), "conv_out (tcgen05)")
            co_sub = None
            if not fuse_out:
                def co_sub(n0, ng, a=a):
                    return self._conv_sub(a, n0, ng, hw[0] * hw[1], "conv_out (tcgen05)", out_f32_of=lambda: self.out_ptr)
            self._emit("conv_out", meta, run_conv_out, sub=co_sub, px=hw[0] * hw[1])
        else:
            co_args = (act.data_ptr(), W["conv_out.w"].data_ptr(), W["conv_out.b"].data_ptr())
            co_tail = (b, c0, hw[0], hw[1], self.cout)
            self._emit("conv_out", meta,
                       lambda st: check(lib.dsg_conv_out(*co_args, self.out_ptr, *co_tail, st), "conv_out"))
        self.n_launches = len(self.ops) + 1  # time_embed is two launches (zero_stats is one fill kernel)

    def run_timed(self, sample: torch.Tensor, t_float: torch.Tensor, out: Optional[torch.Tensor] = None):
        """Eager replay with a CUDA-event pair around every launch; returns [(name, meta, milliseconds)]."""
        dev = sample.device
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.ops) + 1)]
        ops = self.ops
        hooked = []
        stream = torch.cuda.current_stream(dev)
        for i, op in enumerate(ops):
            def h(st, op=op, i=i):
                op(st)
                evs[i + 1].record(stream)
            hooked.append(h)
        self.ops = hooked
        try:
            evs[0].record(stream)
            self.run(sample, t_float, out)
        finally:
            self.ops = ops
        torch.cuda.synchronize(dev)
        return [(n, m, evs[i].elapsed_time(evs[i + 1])) for i, (n, m) in enumerate(self.op_info)]

    def run(self, sample: torch.Tensor, t_float: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if sample.dtype != torch.float32 or not sample.is_contiguous():
            sample = sample.contiguous().float()
        if out is None:
            out = torch.empty((self.b, self.cout, self.h, self.w), dtype=torch.float32, device=sample.device)
        assert t_float.dtype == torch.float32 and t_float.numel() == self.b and t_float.is_cuda
        self.in_ptr = C.c_void_p(sample.data_ptr())
        self.t_ptr = C.c_void_p(t_float.data_ptr())
        self.out_ptr = C.c_void_p(out.data_ptr())
        with torch.cuda.device(sample.device):   # launches go to the tensors' device, whatever the caller's current one
            st = torch.cuda.current_stream(sample.device).cuda_stream
            for op in self.ops:
                op(st)
        return out
