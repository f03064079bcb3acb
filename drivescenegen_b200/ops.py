"""Tensor-level wrappers over the C ABI (include/dsg_b200.h): torch tensors in, torch tensors out.

Used by the parity tests and the micro-benchmarks; the engine calls the C ABI directly with cached pointers.
Every function requires CUDA tensors and raises ``DsgError`` on failure — no fallbacks.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ConvArgs, DsgError, check


def _st(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise DsgError("dsg_b200.ops: CUDA tensors required")


def ddpm_step(eps, sample, noise, table, row: int, row_dev=None, ddim: bool = False):
    _cuda(eps, sample, noise, table)
    lib = _lib.load()
    out = torch.empty_like(sample)
    fn = lib.dsg_ddim_step if ddim else lib.dsg_ddpm_step
    check(fn(eps.data_ptr(), sample.data_ptr(), _p(noise), out.data_ptr(), sample.numel(), table.data_ptr(),
             _p(row_dev), row, _st(sample)), "scheduler step")
    return out


def add_noise(x0, noise, t, sqrt_ac, sqrt_1mac):
    _cuda(x0, noise, t, sqrt_ac, sqrt_1mac)
    lib = _lib.load()
    out = torch.empty_like(x0)
    b = x0.shape[0]
    check(lib.dsg_add_noise(x0.data_ptr(), noise.data_ptr(), t.data_ptr(), sqrt_ac.data_ptr(), sqrt_1mac.data_ptr(),
                            out.data_ptr(), b, x0.numel() // b, _st(x0)), "add_noise")
    return out


def latent_to_image(latent, want_u8=True, want_f32=True):
    _cuda(latent)
    lib = _lib.load()
    n, c, h, w = latent.shape
    u8 = torch.empty((n, h, w, c), dtype=torch.uint8, device=latent.device) if want_u8 else None
    f32 = torch.empty((n, h, w, c), dtype=torch.float32, device=latent.device) if want_f32 else None
    check(lib.dsg_latent_to_image(latent.data_ptr(), _p(u8), _p(f32), n, c, h, w, _st(latent)), "latent_to_image")
    return u8, f32


def time_embed(t, freqs, flip, w1, b1, w2, b2, wp, bp):
    _cuda(t, freqs, w1, b1, w2, b2, wp, bp)
    lib = _lib.load()
    batch, hidden, proj = t.numel(), w1.shape[0], wp.shape[0]
    emb = torch.empty((batch, hidden), dtype=torch.float32, device=t.device)
    out = torch.empty((batch, proj), dtype=torch.float32, device=t.device)
    w1t, w2t = w1.t().contiguous(), w2.t().contiguous()   # the C ABI takes [in][out]
    check(lib.dsg_time_embed(t.data_ptr(), freqs.data_ptr(), freqs.numel(), int(flip), w1t.data_ptr(), b1.data_ptr(),
                             w2t.data_ptr(), b2.data_ptr(), hidden, wp.data_ptr(), bp.data_ptr(), proj, emb.data_ptr(),
                             out.data_ptr(), batch, _st(t)), "time_embed")
    return out, emb


def conv_in(x_nchw, w, b):
    _cuda(x_nchw, w, b)
    lib = _lib.load()
    n, cin, h, wd = x_nchw.shape
    cout = w.shape[0]
    out = torch.empty((n, h, wd, cout), dtype=torch.float16, device=x_nchw.device)
    check(lib.dsg_conv_in(x_nchw.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, cin, h, wd, cout,
                          _st(x_nchw)), "conv_in")
    return out


def conv_out(x_nhwc, w, b):
    _cuda(x_nhwc, w, b)
    lib = _lib.load()
    n, h, wd, cin = x_nhwc.shape
    cout = w.shape[0]
    out = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x_nhwc.device)
    check(lib.dsg_conv_out(x_nhwc.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, cin, h, wd, cout,
                           _st(x_nhwc)), "conv_out")
    return out


def gn_stats(x):
    """Per-channel GroupNorm totals of x fp16 NHWC [n,h,w,c]: int64 [n,c,2] = {sum * 2^24, sum of squares * 2^20}."""
    _cuda(x)
    lib = _lib.load()
    n, h, w, c = x.shape
    st = torch.zeros((n, c, 2), dtype=torch.int64, device=x.device)
    check(lib.dsg_gn_stats(x.data_ptr(), c, st.data_ptr(), n, h * w, _st(x)), "gn_stats")
    return st


def group_norm(x1, x2, gamma, beta, groups: int, eps: float, act: int, stats1=None, stats2=None):
    """x1 [n,h,w,c1] (+ x2 [n,h,w,c2]) fp16 NHWC -> fp16 [n,h,w,c1+c2].  stats: see gn_stats / conv(out_stats=)."""
    _cuda(x1, x2, gamma, beta)
    lib = _lib.load()
    n, h, w, c1 = x1.shape
    c2 = 0 if x2 is None else x2.shape[3]
    hw = h * w
    if stats1 is None:
        stats1 = gn_stats(x1)
    if x2 is not None and stats2 is None:
        stats2 = gn_stats(x2)
    y = torch.empty((n, h, w, c1 + c2), dtype=torch.float16, device=x1.device)
    check(lib.dsg_gn_apply(x1.data_ptr(), c1, stats1.data_ptr(), _p(x2), c2, _p(stats2), gamma.data_ptr(),
                           beta.data_ptr(), eps, act, y.data_ptr(), n, hw, groups, _st(x1)), "gn_apply")
    return y


def pack_conv_weight(mode: int, w, w_sc=None):
    _cuda(w, w_sc)
    lib = _lib.load()
    w = w.float().contiguous()
    cout, cin = w.shape[0], w.shape[1]
    csc = 0
    if w_sc is not None:
        w_sc = w_sc.float().reshape(cout, -1).contiguous()
        csc = w_sc.shape[1]
    k, rows = lib.dsg_packed_k(mode, cin, csc), lib.dsg_packed_rows(mode, cout)
    out = torch.empty((rows, k), dtype=torch.float16, device=w.device)
    check(lib.dsg_pack_conv_weight(mode, w.data_ptr(), cout, cin, _p(w_sc), csc, out.data_ptr(), _st(w)), "pack")
    return out


def conv(mode: int, x, wpacked, cout: int, bias=None, temb=None, temb_off: int = 0, residual=None, sc1=None, sc2=None,
         block_n: int = 0, impl: int = 0, out_stats=None):
    """x fp16 NHWC [n,h,w,cin]; returns fp16 NHWC [n,oh,ow,cout] (see dsg_conv in include/dsg_b200.h).
    out_stats: optional zeroed int64 [n,cout,2] tensor that receives the output's GroupNorm totals."""
    _cuda(x, wpacked, bias, temb, residual, sc1, sc2)
    lib = _lib.load()
    n, h, w, cin = x.shape
    oh, ow = (h // 2, w // 2) if mode == 1 else ((h * 2, w * 2) if mode == 2 else (h, w))
    out = torch.empty((n, oh, ow, cout), dtype=torch.float16, device=x.device)
    a = ConvArgs()
    a.mode, a.n, a.h, a.w, a.cin, a.cout = mode, n, h, w, cin, cout
    a.x = x.data_ptr()
    a.sc1, a.csc1 = _p(sc1), 0 if sc1 is None else sc1.shape[3]
    a.sc2, a.csc2 = _p(sc2), 0 if sc2 is None else sc2.shape[3]
    a.wpacked = wpacked.data_ptr()
    a.bias = _p(bias)
    if temb is not None:
        a.temb, a.temb_stride, a.temb_off = temb.data_ptr(), temb.shape[1], temb_off
    a.residual = _p(residual)
    a.out = out.data_ptr()
    a.block_n, a.impl = block_n, impl
    a.out_stats = _p(out_stats)
    check(lib.dsg_conv(C.byref(a), _st(x)), "dsg_conv")
    return out


def conv_out_tc(x_nhwc, w, b):
    """conv_out on the tensor cores: x fp16 NHWC [n,h,w,cin], w fp32 [cout<=16,cin,3,3], b fp32 [cout] -> fp32 NCHW."""
    _cuda(x_nhwc, w, b)
    lib = _lib.load()
    n, h, wd, cin = x_nhwc.shape
    cout = w.shape[0]
    w16 = torch.zeros((16, cin, 3, 3), dtype=torch.float32, device=w.device)
    w16[:cout] = w
    b16 = torch.zeros(16, dtype=torch.float32, device=w.device)
    b16[:cout] = b
    wp = pack_conv_weight(0, w16)
    out = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x_nhwc.device)
    a = ConvArgs()
    a.mode, a.n, a.h, a.w, a.cin, a.cout = 0, n, h, wd, cin, 16
    a.x, a.wpacked, a.bias = x_nhwc.data_ptr(), wp.data_ptr(), b16.data_ptr()
    a.out_nchw_f32, a.cout_real = out.data_ptr(), cout
    check(lib.dsg_conv(C.byref(a), _st(x_nhwc)), "dsg_conv (conv_out form)")
    return out


def attention(qkv, heads: int, head_dim: int, impl: int = 0, dbg=None):
    """qkv fp16 [n, tokens, 3*heads*head_dim] -> fp16 [n, tokens, heads*head_dim] (see dsg_attention_ex)."""
    _cuda(qkv, dbg)
    lib = _lib.load()
    n, tokens, c3 = qkv.shape
    out = torch.empty((n, tokens, c3 // 3), dtype=torch.float16, device=qkv.device)
    check(lib.dsg_attention_ex(qkv.data_ptr(), out.data_ptr(), n, tokens, heads, head_dim, impl, _p(dbg), _st(qkv)),
          "attention")
    return out
