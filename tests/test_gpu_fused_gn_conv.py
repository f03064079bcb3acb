"""GPU parity of the fused form of dsg_conv: GroupNorm + SiLU (+ the up-block concat) applied to the activation boxes
in shared memory (gn_coef), vs torch fp32 and vs the unfused path (dsg_gn_apply + dsg_conv).

Tolerance: the fused normalisation runs in packed half precision (hfma2 / tanh.f16x2) on fp16 inputs, the unfused one
in fp32 rounded once to fp16: both are within 5e-3 relative L2 of the fp32 reference, and within 3e-3 of each other."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().half()


def _nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


# n, h, w, c1, c2, cout, shortcut, impl (3 = single-CTA halo kernel, 4 = CTA pair)
CASES = [
    (2, 32, 32, 64, 0, 64, False, 4), (2, 32, 32, 64, 0, 64, False, 3),
    (1, 40, 24, 128, 64, 64, True, 4), (2, 24, 16, 128, 64, 128, True, 3),
    (2, 32, 32, 128, 128, 128, True, 4), (1, 32, 32, 256, 128, 256, True, 4),
    (3, 16, 16, 512, 512, 512, True, 4), (2, 24, 8, 256, 0, 256, False, 4), (1, 40, 8, 64, 64, 64, True, 3),
]


@pytest.mark.parametrize("case", CASES)
def test_fused_groupnorm_silu_conv(case):
    from drivescenegen_b200 import ops
    n, h, w, c1, c2, cout, use_sc, impl = case
    g = torch.Generator().manual_seed(31)
    d = _dev()
    c = c1 + c2
    x = (torch.randn(n, c, h, w, generator=g) * 1.3 + 0.2).half().float().to(d)
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).to(d)
    beta = (0.2 * torch.randn(c, generator=g)).to(d)
    wt = (torch.randn(cout, c, 3, 3, generator=g) / math.sqrt(9 * c)).half().float().to(d)
    wsc = (torch.randn(cout, c, 1, 1, generator=g) / math.sqrt(c)).half().float().to(d) if use_sc else None
    bias = torch.randn(cout, generator=g).to(d)
    ref = F.conv2d(F.silu(F.group_norm(x, 32, gamma, beta, 1e-5)), wt, bias, padding=1)
    if use_sc:
        ref = ref + F.conv2d(x, wsc)
    xh = _nhwc(x)
    x1 = xh[..., :c1].contiguous()
    x2 = xh[..., c1:].contiguous() if c2 else None
    st1 = ops.gn_stats(x1)
    st2 = ops.gn_stats(x2) if c2 else None
    wp = ops.pack_conv_weight(0, wt, wsc)
    kw = dict(bias=bias, sc1=x1 if use_sc else None, sc2=x2 if (use_sc and c2) else None)
    # unfused: GroupNorm+SiLU kernel, then the conv
    act = ops.group_norm(x1, x2, gamma, beta, 32, 1e-5, 1, stats1=st1, stats2=st2)
    plain = ops.conv(0, act, wp, cout, impl=impl, **kw)
    # fused: the conv reads the RAW tensors
    coef = ops.gn_coef(st1, st2, gamma, beta, 32, 1e-5, h * w)
    st_out = torch.zeros((n, cout, 2), dtype=torch.int64, device=d)
    fused = ops.conv(0, x1, wp, cout, impl=impl, x2=x2, gn_coef=coef, out_stats=st_out, **kw)
    assert _rel(_nchw(plain), ref) < 5e-3
    assert _rel(_nchw(fused), ref) < 5e-3, _rel(_nchw(fused), ref)
    assert _rel(fused, plain) < 3e-3, _rel(fused, plain)
    # deterministic, and the output statistics ride along as in the unfused form
    again = ops.conv(0, x1, wp, cout, impl=impl, x2=x2, gn_coef=coef, **kw)
    assert torch.equal(again, fused)
    tot = st_out.view(n, cout // 2, 2, 2)[:, :, 0, 0].double().sum(1) / 2 ** 24   # pair totals sit in the even slots
    assert torch.allclose(tot, fused.double().sum(dim=(1, 2, 3)), rtol=1e-3, atol=1e-1)


def test_fused_form_is_refused_where_unsupported():
    from drivescenegen_b200 import _lib, ops
    from drivescenegen_b200._lib import DsgError
    lib = _lib.load()
    assert lib.dsg_conv_gn_fusable(0, 32, 32, 128, 64, 64) == 1
    assert lib.dsg_conv_gn_fusable(0, 4, 4, 64, 64, 64) == 0      # map too small for the halo-reuse kernels
    assert lib.dsg_conv_gn_fusable(3, 32, 32, 64, 64, 64) == 0     # 1x1
    d = _dev()
    x = torch.randn(1, 4, 4, 64, device=d).half()
    wp = ops.pack_conv_weight(0, torch.randn(64, 64, 3, 3, device=d))
    coef = torch.zeros(1, 64, 2, device=d)
    with pytest.raises(DsgError):
        ops.conv(0, x, wp, 64, gn_coef=coef)


def test_conv_out_form_with_fused_norm():
    """conv_norm_out + SiLU + conv_out in one kernel (BLOCK_N = 16, NCHW fp32 output)."""
    import ctypes as C
    from drivescenegen_b200 import _lib, ops
    from drivescenegen_b200._lib import ConvArgs, check
    g = torch.Generator().manual_seed(32)
    d = _dev()
    n, h, w, c = 2, 32, 48, 64
    x = (torch.randn(n, c, h, w, generator=g) * 1.2).half().float().to(d)
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).to(d)
    beta = (0.2 * torch.randn(c, generator=g)).to(d)
    wt = (torch.randn(3, c, 3, 3, generator=g) / 24).half().float().to(d)
    b = torch.randn(3, generator=g).to(d)
    ref = F.conv2d(F.silu(F.group_norm(x, 32, gamma, beta, 1e-5)), wt, b, padding=1)
    xh = _nhwc(x)
    st = ops.gn_stats(xh)
    coef = ops.gn_coef(st, None, gamma, beta, 32, 1e-5, h * w)
    w16 = torch.zeros(16, c, 3, 3, device=d)
    w16[:3] = wt
    b16 = torch.zeros(16, device=d)
    b16[:3] = b
    wp = ops.pack_conv_weight(0, w16)
    out = torch.empty(n, 3, h, w, device=d)
    a = ConvArgs()
    a.mode, a.n, a.h, a.w, a.cin, a.cout = 0, n, h, w, c, 16
    a.x, a.wpacked, a.bias = xh.data_ptr(), wp.data_ptr(), b16.data_ptr()
    a.out_nchw_f32, a.cout_real = out.data_ptr(), 3
    a.cin1, a.gn_coef = c, coef.data_ptr()
    check(_lib.load().dsg_conv(C.byref(a), torch.cuda.current_stream().cuda_stream), "conv_out fused")
    assert _rel(out, ref) < 5e-3, _rel(out, ref)
