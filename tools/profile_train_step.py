#!/usr/bin/env python
"""One eager training forward + backward of the bench workload (B=32, 256x256) inside a cudaProfilerStart/Stop range.

    ncu --profile-from-start off --set full --clock-control none -k regex:"wgrad_kernel|gn_bwd" -c 40 -o ... \
        python tools/profile_train_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from bench import REF_CFG  # noqa: E402
from drivescenegen_b200.hostapi import UNet2DModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = UNet2DModel(sample_size=(args.size, args.size), **REF_CFG).to(dev).train()
x = torch.randn(args.batch, 3, args.size, args.size, device=dev)
noise = torch.randn_like(x)
t = torch.randint(0, 1000, (args.batch,), device=dev)


def step():
    model.zero_grad(set_to_none=True)
    loss = F.mse_loss(model(x, t, return_dict=False)[0], noise) * 65536.0
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
