#!/usr/bin/env python
"""One GroupNorm(+SiLU) backward call at a training shape inside a cudaProfilerStart/Stop range (target of ncu).

    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_bwd \
        -o gpurun_out/prof_gn_bwd python tools/gn_bwd_profile.py --side 256 --c1 64
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drivescenegen_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--side", type=int, default=256)
ap.add_argument("--c1", type=int, default=64)
ap.add_argument("--c2", type=int, default=0)
ap.add_argument("--addend", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, side, c1, c2 = args.batch, args.side, args.c1, args.c2
c = c1 + c2
x1 = torch.randn(B, side, side, c1, device=dev).half()
x2 = torch.randn(B, side, side, c2, device=dev).half() if c2 else None
dy = torch.randn(B, side, side, c, device=dev).half()
addend = torch.randn(B, side, side, c, device=dev).half() if args.addend else None
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
st1 = ops.gn_stats(x1)
st2 = ops.gn_stats(x2) if c2 else None
dx1, dx2 = torch.empty_like(x1), (torch.empty_like(x2) if c2 else None)


def call():
    ops.gn_bwd(dy, x1, x2, gamma, beta, 32, 1e-5, 1, stats1=st1, stats2=st2, addend=addend, dx1=dx1, dx2=dx2,
               want_colsum=True, want_osum=True)


call()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
call()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
