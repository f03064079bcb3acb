#!/usr/bin/env python
"""The optimizer phase of the training step, piece by piece: fused grad-norm, fused AdamW, the re-pack of the fp16 conv
weights (forward + data-gradient packings), CUDA-event medians."""
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
from drivescenegen_b200 import ops  # noqa: E402
from drivescenegen_b200.hostapi import Accelerator, DDPMScheduler, UNet2DModel  # noqa: E402

B, S = int(os.environ.get("B", "4")), 256
dev = torch.device("cuda", 0)
acc = Accelerator(mixed_precision="fp16", gradient_accumulation_steps=1)
torch.manual_seed(0)
model = UNet2DModel(sample_size=(S, S), **REF_CFG).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
model, opt = acc.prepare(model, opt)
sched = DDPMScheduler()
x = (torch.rand(B, 3, S, S) * 2 - 1).to(dev)
noise = torch.randn(B, 3, S, S).to(dev)
t = torch.randint(0, 1000, (B,)).to(dev)
for _ in range(2):
    with acc.accumulate(model):
        loss = F.mse_loss(model(sched.add_noise(x, noise, t), t, return_dict=False)[0], noise)
        acc.backward(loss)
        acc.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
flat_p, fg = model._flat_params, model._flat_grads
flat_g = fg.flat[0]
m, v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
eng = model.engine(train=True)


def timed(fn, n=10):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return round(statistics.median(ts), 4)


out = {"params": flat_p.numel()}
out["grad_norm_ms"] = timed(lambda: ops.grad_norm(flat_g, 1.0, 1.0))
ctl = ops.grad_norm(flat_g, 1.0, 1.0)
out["adamw_ms"] = timed(lambda: ops.adamw_step(flat_p, flat_g, m, v, 1e-5, 0.9, 0.999, 1e-8, 0.01, 1, ctl))
out["repack_ms"] = timed(lambda: eng._run_pack_jobs())
out["pack_jobs"] = len(eng._jobs)
out["packed_halves"] = sum(j[7] * j[8] for j in eng._jobs)
print(json.dumps(out))
