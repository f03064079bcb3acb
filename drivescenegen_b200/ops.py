"""Tensor-level wrappers over the C ABI (include/dsg_b200.h): torch tensors in, torch tensors out.

Used by the parity tests and the micro-benchmarks; the engine calls the C ABI directly with cached pointers.
Every function requires CUDA tensors and raises ``DsgError`` on failure — no fallbacks.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import ConvArgs, DsgError, WgradArgs, check


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
                            sqrt_ac.numel(), out.data_ptr(), b, x0.numel() // b, _st(x0)), "add_noise")
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


def conv_in_stats(x_nchw, w, b):
    """conv_in + the per-channel GroupNorm totals of its output (int64 [n,cout,2], see gn_stats)."""
    _cuda(x_nchw, w, b)
    lib = _lib.load()
    n, cin, h, wd = x_nchw.shape
    cout = w.shape[0]
    out = torch.empty((n, h, wd, cout), dtype=torch.float16, device=x_nchw.device)
    st = torch.zeros((n, cout, 2), dtype=torch.int64, device=x_nchw.device)
    check(lib.dsg_conv_in_stats(x_nchw.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), st.data_ptr(), n, cin, h,
                                wd, cout, _st(x_nchw)), "conv_in_stats")
    return out, st


def conv_out(x_nhwc, w, b):
    _cuda(x_nhwc, w, b)
    lib = _lib.load()
    n, h, wd, cin = x_nhwc.shape
    cout = w.shape[0]
    out = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x_nhwc.device)
    check(lib.dsg_conv_out(x_nhwc.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, cin, h, wd, cout,
                           _st(x_nhwc)), "conv_out")
    return out


def conv_out_fused(x_nhwc, coef, w, b):
    """conv_norm_out + SiLU + conv_out in one pass: x raw fp16 NHWC [n,h,w,64], coef fp32 [n,64,2] from gn_coef (or None
    when x is already activated) -> fp32 NCHW [n,cout,h,w]."""
    _cuda(x_nhwc, coef, w, b)
    lib = _lib.load()
    n, h, wd, cin = x_nhwc.shape
    cout = w.shape[0]
    out = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x_nhwc.device)
    check(lib.dsg_conv_out_fused(x_nhwc.data_ptr(), _p(coef), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, cin, h, wd,
                                 cout, _st(x_nhwc)), "conv_out_fused")
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


def pack_conv_weights_batched(jobs):
    """jobs: [(mode, w fp32 OIHW on the device, w_sc or None), ...] -> list of packed h16 [rows, k] tensors, all packed by
    ONE launch of the shared-memory staged kernel (pack_weights.cu)."""
    lib = _lib.load()
    arr = (_lib.PackJob * len(jobs))()
    outs, keep, begin = [], [], 0
    for j, (mode, w, w_sc) in zip(arr, jobs):
        _cuda(w, w_sc)
        w = w.float().contiguous()
        cout, cin = w.shape[0], w.shape[1]
        csc = 0
        if w_sc is not None:
            w_sc = w_sc.float().reshape(cout, -1).contiguous()
            csc = w_sc.shape[1]
        if mode >= 10:
            k, rows = lib.dsg_packed_k_dgrad(mode - 10, cout), lib.dsg_packed_rows_dgrad(mode - 10, cin)
        else:
            k, rows = lib.dsg_packed_k(mode, cin, csc), lib.dsg_packed_rows(mode, cout)
        out = torch.empty((rows, k), dtype=torch.float16, device=w.device)
        nb = int(lib.dsg_pack_job_blocks(mode, cout, cin, csc))
        if nb < 0:
            raise DsgError("pack_conv_weights_batched: layer too wide for the staged kernel")
        j.mode, j.cout, j.cin, j.csc = mode, cout, cin, csc
        j.w, j.w_sc, j.out = w.data_ptr(), _p(w_sc), out.data_ptr()
        j.k_total, j.rows, j.chunk_begin = k, rows, begin
        begin += nb
        outs.append(out)
        keep.extend([w, w_sc])
    dev = outs[0].device
    jobs_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    check(lib.dsg_pack_conv_weights_batched(jobs_dev.data_ptr(), len(jobs), begin, _st(outs[0])), "pack (batched)")
    torch.cuda.current_stream(dev).synchronize()   # jobs_dev / keep must outlive the launch
    return outs


def pack_conv_weight(mode: int, w, w_sc=None):
    _cuda(w, w_sc)
    lib = _lib.load()
    w = w.float().contiguous()
    cout, cin = w.shape[0], w.shape[1]
    csc = 0
    if w_sc is not None:
        w_sc = w_sc.float().reshape(cout, -1).contiguous()
        csc = w_sc.shape[1]
    if mode >= 10:
        k, rows = lib.dsg_packed_k_dgrad(mode - 10, cout), lib.dsg_packed_rows_dgrad(mode - 10, cin)
    else:
        k, rows = lib.dsg_packed_k(mode, cin, csc), lib.dsg_packed_rows(mode, cout)
    out = torch.empty((rows, k), dtype=torch.float16, device=w.device)
    check(lib.dsg_pack_conv_weight(mode, w.data_ptr(), cout, cin, _p(w_sc), csc, out.data_ptr(), _st(w)), "pack")
    return out


def gn_coef(stats1, stats2, gamma, beta, groups: int, eps: float, hw: int):
    """coefficients of the fused (in-conv) GroupNorm+SiLU: fp32 [n, c1+c2, 2] = (a/2, b/2)."""
    _cuda(stats1, stats2, gamma, beta)
    lib = _lib.load()
    n, c1 = stats1.shape[0], stats1.shape[1]
    c2 = 0 if stats2 is None else stats2.shape[1]
    coef = torch.empty((n, c1 + c2, 2), dtype=torch.float32, device=stats1.device)
    check(lib.dsg_gn_coef(c1, stats1.data_ptr(), c2, _p(stats2), gamma.data_ptr(), beta.data_ptr(), eps, coef.data_ptr(),
                          n, hw, groups, _st(stats1)), "gn_coef")
    return coef


def conv(mode: int, x, wpacked, cout: int, bias=None, temb=None, temb_off: int = 0, residual=None, sc1=None, sc2=None,
         block_n: int = 0, impl: int = 0, out_stats=None, x2=None, gn_coef=None):
    """x fp16 NHWC [n,h,w,cin]; returns fp16 NHWC [n,oh,ow,cout] (see dsg_conv in include/dsg_b200.h).
    out_stats: optional zeroed int64 [n,cout,2] tensor that receives the output's GroupNorm totals.
    gn_coef (+ optional x2): fused GroupNorm+SiLU of the raw input cat(x, x2)."""
    _cuda(x, wpacked, bias, temb, residual, sc1, sc2, x2, gn_coef)
    lib = _lib.load()
    n, h, w, cin = x.shape
    cin1 = cin
    if x2 is not None:
        cin = cin + x2.shape[3]
    oh, ow = (h // 2, w // 2) if mode in (1, 4) else ((h * 2, w * 2) if mode == 2 else (h, w))
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
    a.x2, a.cin1, a.gn_coef = _p(x2), cin1, _p(gn_coef)
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


# ---------------------------------------------------------------------------------------------------------------
# training path
def grad_scale(dout):
    """power-of-two scale pair {s, 1/s} (device float[2]) with amax(|dout|) * s in [1, 2)."""
    _cuda(dout)
    lib = _lib.load()
    parts = 148 * 4
    partial = torch.empty(parts, dtype=torch.float32, device=dout.device)
    scale = torch.empty(2, dtype=torch.float32, device=dout.device)
    check(lib.dsg_grad_scale(dout.data_ptr(), dout.numel(), partial.data_ptr(), parts, scale.data_ptr(), _st(dout)),
          "grad_scale")
    return scale


def gn_bwd(dy, x1, x2, gamma, beta, groups: int, eps: float, act: int, stats1=None, stats2=None, addend=None,
           dx1=None, dx2=None, acc1=False, acc2=False, want_colsum=False, inv_scale=None, want_osum=False):
    """GroupNorm(+SiLU) backward.  Returns (dx1, dx2, dgamma, dbeta, colsum_per_sample or None) and, with want_osum,
    additionally (column sums of the stored dx1 [c1], of the stored dx2 [c2] or None)."""
    _cuda(dy, x1, x2, gamma, beta, addend)
    lib = _lib.load()
    n, h, w, c1 = x1.shape
    c2 = 0 if x2 is None else x2.shape[3]
    c, hw = c1 + c2, h * w
    if stats1 is None:
        stats1 = gn_stats(x1)
    if x2 is not None and stats2 is None:
        stats2 = gn_stats(x2)
    wave = int(os.environ.get("DSG_GN_BWD_WAVE", 148 * 2))   # ONE wave of the 2-CTA-per-SM kernels (measured: 2 waves 6-18 % slower)
    chunks = max(1, min(64, wave // n, -(-hw // 64)))
    # chunk partials + the per-sample (sum g, sum g xh) slot, then the per-channel dx coefficients [n][c][4]
    partial = torch.empty(n * (chunks + 1) * c * 2 + n * c * 4, dtype=torch.float32, device=dy.device)
    if dx1 is None:
        dx1 = torch.empty_like(x1)
    if x2 is not None and dx2 is None:
        dx2 = torch.empty_like(x2)
    parts = max(1, min(wave // n, -(-hw // 32))) if (want_colsum or want_osum) else 0
    colsum = torch.empty((n, parts, c), dtype=torch.float32, device=dy.device) if want_colsum else None
    osum1 = torch.empty((n, parts, c1), dtype=torch.float32, device=dy.device) if want_osum else None
    osum2 = torch.empty((n, parts, c2), dtype=torch.float32, device=dy.device) if (want_osum and c2) else None
    check(lib.dsg_gn_bwd(dy.data_ptr(), x1.data_ptr(), c1, stats1.data_ptr(), _p(x2), c2, _p(stats2), gamma.data_ptr(),
                         beta.data_ptr(), eps, act, partial.data_ptr(), chunks, _p(addend), dx1.data_ptr(), int(acc1),
                         _p(dx2), int(acc2), _p(colsum), _p(osum1), _p(osum2), parts, n, hw, groups, _st(dy)), "gn_bwd")
    dgamma = torch.empty(c, dtype=torch.float32, device=dy.device)
    dbeta = torch.empty(c, dtype=torch.float32, device=dy.device)
    check(lib.dsg_gn_bwd_params(partial.data_ptr(), n, chunks, c, _p(inv_scale), dgamma.data_ptr(), dbeta.data_ptr(),
                                _st(dy)), "gn_bwd_params")
    per_n = None
    if want_colsum:
        per_n = torch.empty((n, c), dtype=torch.float32, device=dy.device)
        check(lib.dsg_colsum_finalize(colsum.data_ptr(), n, parts, c, per_n.data_ptr(), c, 0, None, None, None,
                                      _st(dy)), "colsum_finalize")
    if want_osum:
        outs = []
        for os_, cw in ((osum1, c1), (osum2, c2)):
            if os_ is None:
                outs.append(None)
                continue
            tot = torch.empty(cw, dtype=torch.float32, device=dy.device)
            check(lib.dsg_colsum_finalize(os_.data_ptr(), n, parts, cw, None, 0, 0, None, tot.data_ptr(), None,
                                          _st(dy)), "colsum_finalize")
            outs.append(tot)
        return dx1, dx2, dgamma, dbeta, per_n, outs[0], outs[1]
    return dx1, dx2, dgamma, dbeta, per_n


def colsum(x, inv_scale=None):
    """sum over all leading dims of an fp16 [..., c] tensor -> fp32 [c]."""
    _cuda(x)
    lib = _lib.load()
    c = x.shape[-1]
    rows = x.numel() // c
    parts = max(1, min(148 * 4, -(-rows // 64)))
    partial = torch.empty((parts, c), dtype=torch.float32, device=x.device)
    out = torch.empty(c, dtype=torch.float32, device=x.device)
    check(lib.dsg_colsum_h16(x.data_ptr(), rows, c, partial.data_ptr(), parts, _st(x)), "colsum_h16")
    check(lib.dsg_colsum_finalize(partial.data_ptr(), 1, parts, c, None, 0, 0, _p(inv_scale), out.data_ptr(), None,
                                  _st(x)), "colsum_finalize")
    return out


def conv_wgrad(mode: int, x, dy, ci_total: int = 0, ci_off: int = 0, grad=None, accumulate=False, inv_scale=None,
               impl: int = 0):
    """weight gradient of a dsg_conv-style conv: x fp16 NHWC (forward input), dy fp16 NHWC (output gradient) ->
    fp32 OIHW [cout, ci_total, 3, 3] (mode 3: [cout, ci_total])."""
    _cuda(x, dy)
    lib = _lib.load()
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    ci_total = ci_total or cin
    if grad is None:
        shape = (cout, ci_total) if mode == 3 else (cout, ci_total, 3, 3)
        grad = torch.zeros(shape, dtype=torch.float32, device=x.device)
    nbytes = max(16, lib.dsg_wgrad_workspace_bytes(mode, n, h, w, cin, cout))
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device)
    a = WgradArgs()
    a.mode, a.n, a.h, a.w, a.cin, a.cout = mode, n, h, w, cin, cout
    a.x, a.dy, a.grad = x.data_ptr(), dy.data_ptr(), grad.data_ptr()
    a.ci_total, a.ci_off, a.accumulate = ci_total, ci_off, int(accumulate)
    a.inv_scale = _p(inv_scale)
    a.workspace, a.workspace_bytes, a.impl = ws.data_ptr(), nbytes, impl
    check(lib.dsg_conv_wgrad(C.byref(a), _st(x)), "conv_wgrad")
    return grad


def conv_out_dgrad(dout_nchw, w, scale=None):
    """data gradient of conv_out: dout fp32 NCHW [n,cout,h,w], w fp32 [cout,cin,3,3] -> fp16 NHWC [n,h,w,cin] (* s)."""
    _cuda(dout_nchw, w)
    lib = _lib.load()
    cout, cin = w.shape[0], w.shape[1]
    wt = torch.empty((cin, cout, 3, 3), dtype=torch.float32, device=w.device)
    check(lib.dsg_conv_out_dgrad_weight(w.contiguous().data_ptr(), cout, cin, None, wt.data_ptr(), _st(w)),
          "conv_out_dgrad_weight")
    zero_b = torch.zeros(cin, dtype=torch.float32, device=w.device)
    if scale is None:
        return conv_in(dout_nchw, wt, zero_b)
    n, _, h, wd = dout_nchw.shape
    out = torch.empty((n, h, wd, cin), dtype=torch.float16, device=w.device)
    check(lib.dsg_conv_in_scaled(dout_nchw.data_ptr(), scale.data_ptr(), wt.data_ptr(), zero_b.data_ptr(), out.data_ptr(),
                                 n, cout, h, wd, cin, _st(w)), "conv_in_scaled")
    return out


def small_wgrad(wide, narrow, conv_out_form: bool, inv_scale=None):
    """conv_in / conv_out weight gradient.  wide fp16 NHWC [n,h,w,wc], narrow fp32 NCHW [n,nc,h,w]."""
    _cuda(wide, narrow)
    lib = _lib.load()
    n, h, w, wc = wide.shape
    nc = narrow.shape[1]
    parts = min(n * h, 148 * 4)
    partial = torch.empty((parts + 1, nc * 9 * wc + nc), dtype=torch.float32, device=wide.device)
    dw = torch.empty((nc, wc, 3, 3) if conv_out_form else (wc, nc, 3, 3), dtype=torch.float32, device=wide.device)
    nsum = torch.empty(nc, dtype=torch.float32, device=wide.device)
    check(lib.dsg_small_wgrad(wide.data_ptr(), narrow.data_ptr(), n, h, w, wc, nc, int(conv_out_form),
                              partial.data_ptr(), parts, _p(inv_scale), dw.data_ptr(), nsum.data_ptr(), _st(wide)),
          "small_wgrad")
    return dw, nsum


def attention_train(qkv, heads: int, head_dim: int):
    """training forward on the tcgen05 kernel: returns (out fp16 [n,tokens,c], lse fp32 [n,heads,tokens])."""
    _cuda(qkv)
    lib = _lib.load()
    n, tokens, c3 = qkv.shape
    out = torch.empty((n, tokens, c3 // 3), dtype=torch.float16, device=qkv.device)
    lse = torch.empty((n, heads, tokens), dtype=torch.float32, device=qkv.device)
    check(lib.dsg_attention_train(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), n, tokens, heads, head_dim,
                                  _st(qkv)), "attention_train")
    return out, lse


def attention_bwd(qkv, out, dout, heads: int, head_dim: int, lse=None):
    """lse given (attention_train): tcgen05 backward; else the CUDA-core kernels."""
    _cuda(qkv, out, dout, lse)
    lib = _lib.load()
    n, tokens, _ = qkv.shape
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(2 * n * heads * tokens, dtype=torch.float32, device=qkv.device)
    check(lib.dsg_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), ws.data_ptr(),
                                _p(lse), n, tokens, heads, head_dim, _st(qkv)), "attention_bwd")
    return dqkv


def lin_dgrad_small(dy, w, pre=None, dy_off: int = 0, rows: int = 0):
    _cuda(dy, w, pre)
    lib = _lib.load()
    rows = rows or w.shape[0]
    cols = w.shape[1]
    batch = dy.shape[0]
    dx = torch.empty((batch, cols), dtype=torch.float32, device=dy.device)
    check(lib.dsg_lin_dgrad_small(dy.data_ptr(), dy.shape[1], dy_off, w.data_ptr(), rows, cols, _p(pre), dx.data_ptr(),
                                  batch, _st(dy)), "lin_dgrad_small")
    return dx


def lin_wgrad_small(dy, x, rows: int = 0, dy_off: int = 0, inv_scale=None):
    _cuda(dy, x)
    lib = _lib.load()
    rows = rows or dy.shape[1]
    cols, batch = x.shape[1], dy.shape[0]
    dw = torch.empty((rows, cols), dtype=torch.float32, device=dy.device)
    db = torch.empty(rows, dtype=torch.float32, device=dy.device)
    check(lib.dsg_lin_wgrad_small(dy.data_ptr(), dy.shape[1], dy_off, x.data_ptr(), x.shape[1], rows, cols, batch,
                                  _p(inv_scale), dw.data_ptr(), db.data_ptr(), _st(dy)), "lin_wgrad_small")
    return dw, db


def grad_norm(g, inv_loss_scale: float = 1.0, max_norm: float = 0.0):
    """device float[3] = {norm, unscale-and-clip coefficient, nonfinite flag} of a flat fp32 gradient buffer."""
    _cuda(g)
    lib = _lib.load()
    parts = 148 * 8
    partial = torch.empty(parts, dtype=torch.float64, device=g.device)
    out = torch.empty(3, dtype=torch.float32, device=g.device)
    check(lib.dsg_grad_norm(g.data_ptr(), g.numel(), partial.data_ptr(), parts, inv_loss_scale, max_norm,
                            out.data_ptr(), _st(g)), "grad_norm")
    return out


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step: int, ctl=None):
    _cuda(p, g, m, v, ctl)
    lib = _lib.load()
    check(lib.dsg_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps,
                             weight_decay, step, _p(ctl), _st(p)), "adamw_step")
