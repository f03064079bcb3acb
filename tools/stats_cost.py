#!/usr/bin/env python
"""What do the fused GroupNorm statistics cost the conv epilogue?  Replays every conv launch of the sampling program with
and without out_stats (default implementation), CUDA events, L2 flushed.  Measurement tool only.

    python tools/stats_cost.py [--batch 16] [--size 256]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters", type=int, default=7)
    args = ap.parse_args()
    from bench import REF_CFG
    from drivescenegen_b200 import _lib
    from drivescenegen_b200._lib import ConvArgs
    from drivescenegen_b200.hostapi import UNet2DModel
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    torch.manual_seed(0)
    B, S = args.batch, args.size
    model = UNet2DModel(sample_size=(S, S), **REF_CFG).to(dev).eval()
    eng = model.engine()
    eng.l2_group = 0
    prog = eng.program(B, S, S)
    prog.run(torch.randn(B, 3, S, S, device=dev), torch.full((B,), 500.0, device=dev))
    torch.cuda.synchronize()
    st = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    convs = [(meta, a) for (name, meta), a in zip([oi for oi in prog.op_info if oi[0] == "conv"],
                                                   [k for k in prog.keep if isinstance(k, ConvArgs)])]
    tot = [0.0, 0.0]

    def timed(a):
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.dsg_conv(C.byref(a), st)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        return ms[len(ms) // 2]

    for meta, a in convs:
        keep = a.out_stats
        if not keep:
            continue
        t_with = timed(a)
        a.out_stats = None
        t_without = timed(a)
        a.out_stats = keep
        tot[0] += t_with
        tot[1] += t_without
        print(f"mode{a.mode} {a.h:3d}x{a.w:<3d} cin{a.cin:5d}+{a.csc1 + a.csc2:<4d} cout{a.cout:4d} res{int(bool(a.residual))} "
              f"with stats {t_with * 1e3:7.1f} us  without {t_without * 1e3:7.1f} us  (+{(t_with / t_without - 1) * 100:4.1f} %)",
              flush=True)
    print(f"total with {tot[0]:.3f} ms, without {tot[1]:.3f} ms")


if __name__ == "__main__":
    main()
