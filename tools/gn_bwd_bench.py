#!/usr/bin/env python
"""GroupNorm(+SiLU) backward at the training step's shapes (B = 32, 256x256 model): per-call time and GB/s of the
algorithmic 3 passes (read x, read dy, write dx).  `DSG_LIB=path` loads another build of libdsg_b200 for A/B runs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drivescenegen_b200 import _lib, ops  # noqa: E402

if os.environ.get("DSG_LIB"):
    _lib.LIB_PATH = os.environ["DSG_LIB"]
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", "32"))
# (hw side, c1, c2, addend, colsum/osum) of the reference U-Net's GroupNorms at 256x256
SHAPES = [(256, 64, 0, False), (256, 64, 0, True), (256, 128, 64, True), (256, 64, 64, True), (128, 128, 0, True),
          (128, 256, 128, True), (64, 256, 0, True), (64, 512, 256, True), (32, 512, 0, True), (32, 512, 512, True)]
out = []
for side, c1, c2, add in SHAPES:
    c = c1 + c2
    x1 = torch.randn(B, side, side, c1, device=dev).half()
    x2 = torch.randn(B, side, side, c2, device=dev).half() if c2 else None
    dy = torch.randn(B, side, side, c, device=dev).half()
    addend = torch.randn(B, side, side, c, device=dev).half() if add else None
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    st1 = ops.gn_stats(x1)
    st2 = ops.gn_stats(x2) if c2 else None
    dx1, dx2 = torch.empty_like(x1), (torch.empty_like(x2) if c2 else None)

    def call():
        ops.gn_bwd(dy, x1, x2, gamma, beta, 32, 1e-5, 1, stats1=st1, stats2=st2, addend=addend, dx1=dx1, dx2=dx2,
                   want_colsum=True, want_osum=True)
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    alg = 3 * B * side * side * c * 2
    out.append({"side": side, "c1": c1, "c2": c2, "addend": add, "ms": ms, "alg_gbs": alg / ms / 1e6})
    print(f"{side:4d} c={c1}+{c2} add={int(add)}  {ms * 1e3:8.1f} us   {alg / ms / 1e6:7.0f} GB/s (3-pass algorithmic)")
print(json.dumps({"lib": _lib.LIB_PATH, "rows": out}))
