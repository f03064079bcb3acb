"""GPU parity of the training-path kernels (backward + optimizer) against torch fp32 autograd, through the C ABI.

Tolerances: operands / activation gradients are fp16 with fp32 accumulation, so relative L2 error <= 5e-3 against an
fp32 reference evaluated on the SAME fp16-rounded inputs (written next to each assert).
"""
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


def _nhwc(x):  # NCHW fp32 -> NHWC fp16
    return x.permute(0, 2, 3, 1).contiguous().half()


def _nchw(x):  # NHWC -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def _fwd_ref(mode, x, w):
    if mode == 0:
        return F.conv2d(x, w, padding=1)
    if mode == 1:
        return F.conv2d(x, w, stride=2, padding=1)
    if mode == 2:
        return F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    return F.conv2d(x, w)


# mode, n, h, w, cin, cout
DGRAD_CASES = [
    (0, 2, 32, 32, 64, 64), (0, 1, 24, 40, 192, 64), (0, 2, 16, 16, 128, 256), (0, 1, 32, 32, 64, 384),
    (1, 2, 32, 32, 64, 64), (1, 1, 64, 32, 128, 128),
    (2, 2, 16, 16, 64, 64), (2, 1, 32, 16, 128, 128),
    (3, 2, 16, 16, 128, 384), (3, 1, 32, 32, 512, 512),
]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_conv_dgrad_matches_autograd(case):
    """input gradient of every conv mode = a dsg_conv over dy with the mode-1x packed weights."""
    from drivescenegen_b200 import ops
    mode, n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(11)
    d = _dev()
    k = 1 if mode == 3 else 3
    x = torch.randn(n, cin, h, w, generator=g).half().float().to(d).requires_grad_(True)
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).half().float().to(d)
    y = _fwd_ref(mode, x, wt)
    dy = torch.randn(y.shape, generator=g).half().float().to(d)
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    wp = ops.pack_conv_weight(10 + mode, wt)
    run_mode = {0: 0, 1: 2, 2: 4, 3: 3}[mode]
    for impl in (0, 1):
        dx = ops.conv(run_mode, _nhwc(dy), wp, cin, impl=impl)
        assert dx.shape == (n, h, w, cin)
        err = _rel(_nchw(dx), dx_ref)
        assert err < 5e-3, f"dgrad mode {mode} impl {impl}: rel {err}"


# mode, n, h, w, cin, cout
WGRAD_CASES = [
    (0, 2, 32, 32, 64, 64), (0, 3, 24, 40, 192, 64), (0, 2, 16, 16, 128, 256), (0, 1, 32, 32, 256, 128),
    (0, 1, 24, 8, 64, 128),
    (1, 2, 32, 32, 64, 64), (1, 1, 64, 32, 128, 128),
    (2, 2, 16, 16, 64, 64), (2, 1, 32, 16, 128, 256),
    (3, 2, 16, 16, 128, 384), (3, 1, 32, 32, 512, 512), (3, 2, 16, 16, 64, 64), (3, 1, 16, 16, 192, 128),
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_conv_wgrad_matches_autograd(case):
    from drivescenegen_b200 import ops
    mode, n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(12)
    d = _dev()
    k = 1 if mode == 3 else 3
    x = torch.randn(n, cin, h, w, generator=g).half().float().to(d)
    wt = torch.zeros(cout, cin, k, k, device=d, requires_grad=True)
    y = _fwd_ref(mode, x, wt)
    dy = torch.randn(y.shape, generator=g).half().float().to(d)
    (dw_ref,) = torch.autograd.grad(y, wt, dy)
    inv = torch.tensor([0.5], device=d)
    for impl in (1, 2):
        dw = ops.conv_wgrad(mode, _nhwc(x), _nhwc(dy), impl=impl)
        ref = dw_ref.reshape(dw.shape)
        err = _rel(dw, ref)
        assert err < 2e-3, f"wgrad mode {mode} impl {impl}: rel {err}"
    # column offset into a wider gradient (conv_shortcut over cat(x1, x2)), inverse scale, accumulate
    tot = cin + 64
    grad = torch.ones((cout, tot) if mode == 3 else (cout, tot, 3, 3), device=d)
    ops.conv_wgrad(mode, _nhwc(x), _nhwc(dy), ci_total=tot, ci_off=64, grad=grad, accumulate=True, inv_scale=inv)
    want = torch.ones_like(grad)
    want[:, 64:] += 0.5 * dw_ref.reshape((cout, cin) if mode == 3 else (cout, cin, 3, 3))
    assert _rel(grad, want) < 2e-3
    # deterministic: bit-identical on a second run
    a = ops.conv_wgrad(mode, _nhwc(x), _nhwc(dy))
    b = ops.conv_wgrad(mode, _nhwc(x), _nhwc(dy))
    assert torch.equal(a, b)


def test_conv_wgrad_small_map_uses_cuda_core_path():
    from drivescenegen_b200 import ops
    from drivescenegen_b200._lib import DsgError
    d = _dev()
    x = torch.randn(1, 4, 4, 64, device=d).half()
    dy = torch.randn(1, 4, 4, 64, device=d).half()
    ops.conv_wgrad(0, x, dy)  # auto: naive kernel
    with pytest.raises(DsgError):
        ops.conv_wgrad(0, x, dy, impl=2)


GN_CASES = [
    # n, h, w, c1, c2, act, addend, acc
    (2, 16, 16, 64, 0, 1, False, False),
    (2, 24, 8, 128, 64, 1, True, True),
    (1, 32, 32, 256, 128, 1, True, False),
    (3, 8, 8, 512, 512, 1, False, True),
    (2, 16, 16, 512, 0, 0, True, False),
    (2, 64, 64, 64, 64, 1, True, True),
]


@pytest.mark.parametrize("case", GN_CASES)
def test_gn_bwd_matches_autograd(case):
    from drivescenegen_b200 import ops
    n, h, w, c1, c2, act, use_add, acc = case
    g = torch.Generator().manual_seed(13)
    d = _dev()
    c = c1 + c2
    x = (torch.randn(n, c, h, w, generator=g) * 1.5 + 0.3).half().float().to(d).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).to(d).requires_grad_(True)
    beta = (0.2 * torch.randn(c, generator=g)).to(d).requires_grad_(True)
    y = F.group_norm(x, 32, gamma, beta, 1e-5)
    if act:
        y = F.silu(y)
    dy = torch.randn(y.shape, generator=g).half().float().to(d)
    dx_ref, dg_ref, db_ref = torch.autograd.grad(y, (x, gamma, beta), dy)
    xh = _nhwc(x.detach())
    x1 = xh[..., :c1].contiguous()
    x2 = xh[..., c1:].contiguous() if c2 else None
    addend = torch.randn(n, h, w, c, generator=g).half().to(d) if use_add else None
    old1 = torch.randn(n, h, w, c1, generator=g).half().to(d)
    old2 = torch.randn(n, h, w, c2, generator=g).half().to(d) if c2 else None
    dx1 = old1.clone() if acc else None
    dx2 = old2.clone() if (acc and c2) else None
    inv = torch.tensor([0.25], device=d)
    dx1, dx2, dgamma, dbeta, per_n, os1, os2 = ops.gn_bwd(_nhwc(dy), x1, x2, gamma.detach(), beta.detach(), 32, 1e-5,
                                                          act, addend=addend, dx1=dx1, dx2=dx2, acc1=acc, acc2=acc,
                                                          want_colsum=True, inv_scale=inv, want_osum=True)
    want = _nhwc(dx_ref).float()
    colsum_ref = want.sum(dim=(1, 2))
    if use_add:
        want = want + addend.float()
    if acc:
        want = want + torch.cat([old1, old2], -1).float() if c2 else want + old1.float()
    got = torch.cat([dx1, dx2], -1) if c2 else dx1
    assert _rel(got, want) < 3e-3, _rel(got, want)   # fp16 storage of dx: ~1e-3
    assert _rel(dgamma, 0.25 * dg_ref) < 3e-3
    assert _rel(dbeta, 0.25 * db_ref) < 3e-3
    assert _rel(per_n, colsum_ref) < 5e-3
    # column sums of the stored gradients (the producer's bias gradient): exactly what a reader of dx would sum
    assert _rel(os1, dx1.float().sum(dim=(0, 1, 2))) < 1e-4
    if c2:
        assert _rel(os2, dx2.float().sum(dim=(0, 1, 2))) < 1e-4


@pytest.mark.parametrize("shape", [(2, 256, 16), (1, 1024, 64), (2, 200, 8)])
def test_attention_bwd_matches_autograd(shape):
    from drivescenegen_b200 import ops
    n, tokens, heads = shape
    hd, c = 8, shape[2] * 8
    g = torch.Generator().manual_seed(14)
    d = _dev()
    qkv = torch.randn(n, tokens, 3 * c, generator=g).half().to(d)
    q, k, v = [t.float().reshape(n, tokens, heads, hd).transpose(1, 2).requires_grad_(True)
               for t in qkv.split(c, dim=-1)]
    o = F.scaled_dot_product_attention(q, k, v)
    do = torch.randn(n, tokens, c, generator=g).half().to(d)
    dq, dk, dv = torch.autograd.grad(o, (q, k, v), do.float().reshape(n, tokens, heads, hd).transpose(1, 2))
    ref = torch.cat([t.transpose(1, 2).reshape(n, tokens, c) for t in (dq, dk, dv)], -1)
    o16 = o.detach().transpose(1, 2).reshape(n, tokens, c).half()
    got = ops.attention_bwd(qkv, o16, do, heads, hd)
    for i, nm in enumerate("qkv"):
        err = _rel(got[..., i * c:(i + 1) * c], ref[..., i * c:(i + 1) * c])
        assert err < 5e-3, f"d{nm}: rel {err}"
    if tokens % 128 == 0:
        # tcgen05 path: forward with log-sum-exp, backward from it (fp16 P / dS operands: ~1e-3 each)
        o_tc, lse = ops.attention_train(qkv, heads, hd)
        assert _rel(o_tc, o16) < 3e-3
        s = torch.einsum("bhid,bhjd->bhij", q.detach(), k.detach()) / math.sqrt(hd)
        lse_ref = torch.logsumexp(s, -1) / math.log(2.0)
        assert (lse - lse_ref).abs().max().item() < 2e-3
        got = ops.attention_bwd(qkv, o_tc, do, heads, hd, lse=lse)
        for i, nm in enumerate("qkv"):
            err = _rel(got[..., i * c:(i + 1) * c], ref[..., i * c:(i + 1) * c])
            assert err < 1e-2, f"tcgen05 d{nm}: rel {err}"


def test_conv_out_and_conv_in_backward():
    from drivescenegen_b200 import ops
    g = torch.Generator().manual_seed(15)
    d = _dev()
    n, h, w, c0 = 2, 32, 24, 64
    # conv_out: act [n,64,h,w] -> out [n,3,h,w]
    act = torch.randn(n, c0, h, w, generator=g).half().float().to(d).requires_grad_(True)
    wo = (torch.randn(3, c0, 3, 3, generator=g) / 24).to(d).requires_grad_(True)
    bo = torch.zeros(3, device=d, requires_grad=True)
    out = F.conv2d(act, wo, bo, padding=1)
    dout = (torch.randn(out.shape, generator=g) * 3e-5).to(d)   # small, like an unscaled MSE gradient
    dact_ref, dwo_ref, dbo_ref = torch.autograd.grad(out, (act, wo, bo), dout)
    scale = ops.grad_scale(dout)
    s = scale[0].item()
    amax = dout.abs().max().item()
    assert 1.0 <= amax * s < 2.0 and abs(scale[1].item() * s - 1.0) < 1e-6 and math.log2(s) == int(math.log2(s))
    dact = ops.conv_out_dgrad(dout, wo.detach(), scale)
    assert _rel(_nchw(dact) / s, dact_ref) < 3e-3
    dwo, dbo = ops.small_wgrad(_nhwc(act.detach()), dout, True)
    assert _rel(dwo, dwo_ref) < 1e-3 and _rel(dbo, dbo_ref) < 1e-4
    # conv_in: x [n,3,h,w] fp32 -> y [n,64,h,w]
    x = torch.randn(n, 3, h, w, generator=g).to(d)
    wi = torch.zeros(c0, 3, 3, 3, device=d, requires_grad=True)
    y = F.conv2d(x, wi, padding=1)
    dy = torch.randn(y.shape, generator=g).half().float().to(d)
    (dwi_ref,) = torch.autograd.grad(y, wi, dy)
    dwi, _ = ops.small_wgrad(_nhwc(dy), x, False, inv_scale=scale[1:])
    assert _rel(dwi, dwi_ref / s) < 1e-3
    assert _rel(ops.colsum(_nhwc(dy)), dy.sum(dim=(0, 2, 3))) < 1e-3


def test_grad_scale_handles_zero_and_nonfinite():
    from drivescenegen_b200 import ops
    d = _dev()
    assert ops.grad_scale(torch.zeros(1000, device=d))[0].item() == 1.0
    x = torch.randn(1000, device=d)
    x[17] = float("inf")
    assert ops.grad_scale(x)[0].item() == 1.0
    x[17] = float("nan")
    assert ops.grad_scale(x)[0].item() == 1.0


def test_small_linear_backward():
    from drivescenegen_b200 import ops
    g = torch.Generator().manual_seed(16)
    d = _dev()
    batch, rows, cols = 5, 300, 256
    x = torch.randn(batch, cols, generator=g).to(d).requires_grad_(True)
    wt = torch.randn(rows, cols, generator=g).to(d).requires_grad_(True)
    b = torch.zeros(rows, device=d, requires_grad=True)
    y = F.linear(F.silu(x), wt, b)
    dy = torch.randn(batch, rows + 7, generator=g).to(d)
    dx_ref, dw_ref, db_ref = torch.autograd.grad(y, (x, wt, b), dy[:, 7:].contiguous())
    dx = ops.lin_dgrad_small(dy, wt.detach(), pre=x.detach(), dy_off=7, rows=rows)
    assert _rel(dx, dx_ref) < 1e-5
    dw, db = ops.lin_wgrad_small(dy, F.silu(x.detach()), rows=rows, dy_off=7)
    assert _rel(dw, dw_ref) < 1e-5 and _rel(db, db_ref) < 1e-5


def test_time_embed_saved_activations():
    from drivescenegen_b200 import _lib
    from drivescenegen_b200._lib import check
    d = _dev()
    g = torch.Generator().manual_seed(17)
    batch, half, hidden, proj = 3, 32, 256, 640
    t = torch.tensor([3.0, 500.0, 999.0], device=d)
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(d)
    w1 = (torch.randn(hidden, 2 * half, generator=g) / 8).to(d)
    b1 = torch.randn(hidden, generator=g).to(d) * 0.1
    w2 = (torch.randn(hidden, hidden, generator=g) / 16).to(d)
    b2 = torch.randn(hidden, generator=g).to(d) * 0.1
    wp = (torch.randn(proj, hidden, generator=g) / 16).to(d)
    bp = torch.randn(proj, generator=g).to(d) * 0.1
    emb = torch.empty(batch, hidden, device=d)
    out = torch.empty(batch, proj, device=d)
    saved = torch.empty(batch, 2 * half + 3 * hidden, device=d)
    st = torch.cuda.current_stream().cuda_stream
    check(_lib.load().dsg_time_embed_ex(t.data_ptr(), freqs.data_ptr(), half, 1, w1.t().contiguous().data_ptr(),
                                        b1.data_ptr(), w2.t().contiguous().data_ptr(), b2.data_ptr(), hidden,
                                        wp.data_ptr(), bp.data_ptr(), proj, emb.data_ptr(), out.data_ptr(), batch,
                                        saved.data_ptr(), st), "time_embed_ex")
    arg = t[:, None] * freqs[None]
    e = torch.cat([torch.cos(arg), torch.sin(arg)], -1)
    pre1 = F.linear(e, w1, b1)
    pre2 = F.linear(F.silu(pre1), w2, b2)
    flat = saved.reshape(-1)
    se = flat[:batch * 2 * half].reshape(batch, 2 * half)
    sh = flat[batch * 2 * half:].reshape(3, batch, hidden)
    assert torch.allclose(se, e, atol=2e-4)   # sinf/cosf of arguments up to ~1e3
    assert torch.allclose(sh[0], pre1, atol=1e-3)
    assert torch.allclose(sh[1], F.silu(pre1), atol=1e-3)
    assert torch.allclose(sh[2], pre2, atol=1e-3)
    assert torch.allclose(out, F.linear(F.silu(pre2), wp, bp), atol=1e-3)


@pytest.mark.parametrize("numel", [1000, 1 << 20, (1 << 20) + 3])
def test_grad_norm_and_adamw_match_torch(numel):
    from drivescenegen_b200 import ops
    g = torch.Generator().manual_seed(18)
    d = _dev()
    p0 = torch.randn(numel, generator=g).to(d)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    loss_scale = 1024.0
    for step in range(1, 4):
        grad = torch.randn(numel, generator=g).to(d) * 0.01
        ref_p.grad = grad.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        ctl = ops.grad_norm(grad * loss_scale, 1.0 / loss_scale, 1.0)
        assert abs(ctl[0].item() - norm_ref.item()) <= 1e-5 * norm_ref.item()
        assert ctl[2].item() == 0.0
        ops.adamw_step(p, grad * loss_scale, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, ctl)
        assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=1e-6), (p - ref_p.detach()).abs().max()
    # non-finite gradients: flagged, update skipped (GradScaler semantics)
    bad = torch.randn(numel, generator=g).to(d)
    bad[3] = float("inf")
    ctl = ops.grad_norm(bad, 1.0, 1.0)
    assert ctl[2].item() == 1.0
    before = p.clone()
    ops.adamw_step(p, bad, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-2, 4, ctl)
    assert torch.equal(p, before)
