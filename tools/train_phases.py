#!/usr/bin/env python
"""Where the training step's time goes outside the forward / backward kernels: CUDA-event time per phase of
bench.py's training step (add_noise, forward, loss, backward, clip, optimizer + re-pack, host sync), median of N steps."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from bench import REF_CFG  # noqa: E402
from drivescenegen_b200.hostapi import Accelerator, DDPMScheduler, UNet2DModel, get_cosine_schedule_with_warmup  # noqa: E402

B, S, N = int(os.environ.get("B", "32")), 256, int(os.environ.get("N", "12"))
dev = torch.device("cuda", 0)
acc = Accelerator(mixed_precision="fp16", gradient_accumulation_steps=1)
torch.manual_seed(0)
model = UNet2DModel(sample_size=(S, S), **REF_CFG).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
lr_sched = get_cosine_schedule_with_warmup(optimizer=opt, num_warmup_steps=500, num_training_steps=100000)
model, opt, lr_sched = acc.prepare(model, opt, lr_sched)
sched = DDPMScheduler()
x = (torch.rand(B, 3, S, S) * 2 - 1).to(dev)
noise = torch.randn(B, 3, S, S).to(dev)
t = torch.randint(0, 1000, (B,)).to(dev)
names = ["add_noise", "forward", "loss", "backward", "clip", "opt_step+repack+sync", "sched+zero"]
rows = []
PROFILE = os.environ.get("PROFILE") == "1"   # ncu --profile-from-start off: the kernels of ONE whole step
for it in range(N + 3):
    if PROFILE and it == N + 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    noisy = sched.add_noise(x, noise, t); ev[1].record()
    with acc.accumulate(model):
        pred = model(noisy, t, return_dict=False)[0]; ev[2].record()
        loss = F.mse_loss(pred, noise); ev[3].record()
        acc.backward(loss); ev[4].record()
        acc.clip_grad_norm_(model.parameters(), 1.0); ev[5].record()
        opt.step(); ev[6].record()
        lr_sched.step()
        opt.zero_grad(); ev[7].record()
    torch.cuda.synchronize()
    if it >= 3:
        rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))] + [ev[0].elapsed_time(ev[7])])
if PROFILE:
    torch.cuda.cudart().cudaProfilerStop()
med = [statistics.median(r[i] for r in rows) for i in range(len(names) + 1)]
out = {n: round(m, 3) for n, m in zip(names + ["total"], med)}
print(json.dumps(out))
