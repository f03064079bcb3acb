"""GPU parity: tcgen05 attention kernel (head_dim 8) vs torch, with the kernel's debug taps checked piece by piece."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(qkv, heads, hd):
    n, tokens, c3 = qkv.shape
    c = c3 // 3
    q, k, v = [t.float().view(n, tokens, heads, hd).transpose(1, 2) for t in qkv.split(c, dim=2)]
    return F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n, tokens, c), q, k, v


@pytest.mark.parametrize("n,tokens,heads", [(1, 128, 1), (1, 256, 2), (2, 1024, 64), (1, 4096, 4), (3, 384, 5)])
def test_attention_tc_stages(n, tokens, heads):
    from drivescenegen_b200 import ops
    torch.manual_seed(tokens + heads)
    d = torch.device("cuda", 0)
    hd = 8
    qkv = (torch.randn(n, tokens, 3 * heads * hd) * 1.5).half()
    ref, q, k, v = _ref(qkv, heads, hd)
    dbg = torch.zeros(128 * 128 + 128 * 16, dtype=torch.float32, device=d)
    got = ops.attention(qkv.to(d), heads, hd, impl=2, dbg=dbg)
    torch.cuda.synchronize()
    dbg = dbg.cpu()
    # stage 1: raw scores S = Q K^T of (sample 0, head 0), queries 0..127 x keys 0..127
    s_ref = q[0, 0, :128] @ k[0, 0, :128].T
    s_got = dbg[:128 * 128].view(128, 128)
    assert torch.allclose(s_got, s_ref, atol=2e-2, rtol=1e-3), \
        f"QK^T tile wrong: max err {(s_got - s_ref).abs().max():.4f}; got[0,:4]={s_got[0, :4]}, ref[0,:4]={s_ref[0, :4]}"
    # stage 2: un-normalised output tile; column 8 is the softmax denominator
    o_got = dbg[128 * 128:].view(128, 16)
    s_all = (q[0, 0, :128] @ k[0, 0].T) / math.sqrt(hd)
    p = torch.exp(s_all - s_all.max(dim=1, keepdim=True).values)
    assert torch.allclose(o_got[:, 8], p.sum(dim=1), rtol=5e-3, atol=1e-3), \
        f"row sums wrong: got {o_got[:4, 8]}, ref {p.sum(dim=1)[:4]}"
    o_ref = p @ v[0, 0]
    assert torch.allclose(o_got[:, :8], o_ref, rtol=5e-3, atol=5e-3), \
        f"P V wrong: max err {(o_got[:, :8] - o_ref).abs().max():.4f}; got {o_got[0, :8]}, ref {o_ref[0]}"
    # final result
    err = (got.float().cpu() - ref).abs().max().item()
    assert err < 3e-3, f"attention output max err {err}"


def test_attention_tc_matches_cuda_core_kernel():
    from drivescenegen_b200 import ops
    torch.manual_seed(5)
    d = torch.device("cuda", 0)
    qkv = (torch.randn(2, 1024, 3 * 512) * 2.0).half().to(d)
    a = ops.attention(qkv, 64, 8, impl=2)
    b = ops.attention(qkv, 64, 8, impl=1)
    torch.cuda.synchronize()
    # two fp16-rounded results of the same fp32 math: within one fp16 ulp of each other
    assert torch.allclose(a.float(), b.float(), atol=1e-3, rtol=2e-3), (a.float() - b.float()).abs().max()


def test_attention_tc_rejects_other_shapes():
    from drivescenegen_b200 import ops
    from drivescenegen_b200._lib import DsgError
    d = torch.device("cuda", 0)
    with pytest.raises(DsgError):
        ops.attention(torch.zeros(1, 200, 3 * 64, dtype=torch.float16, device=d), 8, 8, impl=2)   # tokens % 128
    with pytest.raises(DsgError):
        ops.attention(torch.zeros(1, 256, 3 * 64, dtype=torch.float16, device=d), 4, 16, impl=2)  # head_dim 16
