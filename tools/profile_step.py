#!/usr/bin/env python
"""One eager denoise step (U-Net forward + DDPM step) of the bench workload inside a cudaProfilerStart/Stop range.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm \
        -o gpurun_out/prof_conv python tools/profile_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from bench import REF_CFG  # noqa: E402
from drivescenegen_b200 import _lib  # noqa: E402
from drivescenegen_b200._lib import check  # noqa: E402
from drivescenegen_b200.hostapi import DDPMScheduler, UNet2DModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=256)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = UNet2DModel(sample_size=(args.size, args.size), **REF_CFG).to(dev).eval()
sched = DDPMScheduler()
shape = (args.batch, 3, args.size, args.size)
x = torch.randn(shape, device=dev)
z = torch.randn(shape, device=dev)
eps = torch.empty_like(x)
nxt = torch.empty_like(x)
tf = torch.full((args.batch,), 500.0, device=dev)
prog = model.engine().program(args.batch, args.size, args.size)
table = sched.coef_table(dev)
lib = _lib.load()


def step():
    prog.run(x, tf, eps)
    st = torch.cuda.current_stream(dev).cuda_stream
    check(lib.dsg_ddpm_step(eps.data_ptr(), x.data_ptr(), z.data_ptr(), nxt.data_ptr(), x.numel(), table.data_ptr(),
                            None, 500, st), "ddpm_step")


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
