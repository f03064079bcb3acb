#!/usr/bin/env python
"""Small invocations of the kernels added late in round 1, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_new_kernels.py
    compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py

(raster.cu, conv_in_mma_kernel incl. the scaled / partial-tile forms, the three-warpgroup attention kernel, the weight
gradient + its reduce, the small-linear and conv_in/conv_out gradient kernels).  Measurement / hygiene tool only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from drivescenegen_b200 import ops  # noqa: E402
from drivescenegen_b200.hostapi import raster  # noqa: E402

d = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
rng = np.random.default_rng(0)

for shape in [(2, 32, 48, 3), (1, 13, 37, 3), (2, 16, 16, 4)]:
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    raster.gray_masks(img, want_gray3=True)
    raster.image_to_sample(img, channels=3)
raster.agent_threshold(torch.rand(2, 3, 24, 40).to(d))
raster.agent_threshold(torch.rand(1, 3, 7, 9).to(d))

for (n, h, w) in [(2, 16, 32), (1, 13, 37), (3, 9, 16)]:
    x = torch.randn(n, 3, h, w, generator=g).to(d)
    wt = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).to(d)
    b = torch.randn(64, generator=g).to(d)
    ops.conv_in_stats(x, wt, b)
    wo = (torch.randn(3, 64, 3, 3, generator=g) / 24).to(d)
    dout = (torch.randn(n, 3, h, w, generator=g) * 1e-5).to(d)
    ops.conv_out_dgrad(dout, wo, ops.grad_scale(dout))
    ops.small_wgrad(torch.randn(n, h, w, 64, generator=g).half().to(d), dout, True)
    ops.small_wgrad(torch.randn(n, h, w, 64, generator=g).half().to(d), x, False)

for (n, tokens, heads) in [(1, 128, 2), (2, 384, 3), (1, 1024, 4)]:
    qkv = torch.randn(n, tokens, 3 * heads * 8, generator=g).half().to(d)
    ops.attention(qkv, heads, 8)
    ops.attention_train(qkv, heads, 8)

for (mode, cin, cout, hw) in [(0, 64, 64, 16), (0, 128, 256, 16), (3, 64, 128, 16), (2, 64, 64, 8)]:
    x = torch.randn(2, hw, hw, cin, generator=g).half().to(d)
    ohw = hw * 2 if mode == 2 else hw
    dy = torch.randn(2, ohw, ohw, cout, generator=g).half().to(d)
    ops.conv_wgrad(mode, x, dy)

# round 2: GroupNorm backward (statistics / per-sample coefficients / operand-flag-specialised apply variants), the TMA-store
# conv epilogue (64 -> 64 and 1x1 convs), the shared-memory staged weight re-pack, the bilinear resize
for (n, hh, ww, c1, c2, act, add, acc) in [(2, 16, 16, 64, 0, 1, False, False), (2, 24, 8, 128, 64, 1, True, True),
                                           (1, 8, 8, 512, 512, 1, False, True), (2, 16, 16, 256, 0, 0, True, False)]:
    c = c1 + c2
    x1 = torch.randn(n, hh, ww, c1, generator=g).half().to(d)
    x2 = torch.randn(n, hh, ww, c2, generator=g).half().to(d) if c2 else None
    dyh = torch.randn(n, hh, ww, c, generator=g).half().to(d)
    addend = torch.randn(n, hh, ww, c, generator=g).half().to(d) if add else None
    ops.gn_bwd(dyh, x1, x2, torch.ones(c, device=d), torch.zeros(c, device=d), 32, 1e-5, act, addend=addend,
               dx1=torch.zeros_like(x1) if acc else None, dx2=torch.zeros_like(x2) if (acc and c2) else None,
               acc1=acc, acc2=acc and c2 > 0, want_colsum=True, want_osum=True)
for (mode, cin, cout) in [(0, 64, 64), (3, 64, 128), (3, 128, 64)]:
    xh = torch.randn(2, 32, 32, cin, generator=g).half().to(d)
    kk = 1 if mode == 3 else 3
    ops.conv(mode, xh, ops.pack_conv_weight(mode, (torch.randn(cout, cin, kk, kk, generator=g) * 0.05).to(d)), cout,
             bias=torch.zeros(cout, device=d))
jobs = []
for cout, cin in ((64, 64), (72, 40), (136, 200)):
    w3 = torch.randn(cout, cin, 3, 3, generator=g).to(d)
    jobs += [(m, w3, None) for m in (0, 1, 2, 10, 11, 12)]
    jobs += [(0, w3, torch.randn(cout, 48, generator=g).to(d))]
    w1 = torch.randn(cout, cin, 1, 1, generator=g).to(d)
    jobs += [(3, w1, None), (13, w1, None)]
ops.pack_conv_weights_batched(jobs)
raster.image_to_sample(rng.integers(0, 256, (2, 50, 70, 3), dtype=np.uint8), channels=3, size=(32, 48))

dy = torch.randn(4, 600, generator=g).to(d)
wl = torch.randn(600, 256, generator=g).to(d)
ops.lin_dgrad_small(dy, wl)
torch.cuda.synchronize()
print("sanitize_new_kernels: done")
