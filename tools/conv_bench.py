#!/usr/bin/env python
"""Per-layer A/B of the two tcgen05 conv kernels on the reference U-Net's own conv shapes.

    python tools/conv_bench.py [--batch 16] [--size 256] [--iters 10] [--out gpurun_out/conv_bench.json]

Builds the engine program for (batch, size, size), runs one real forward so every buffer holds finite data, then
replays each conv launch with impl = 2 (tap-streaming kernel), impl = 3 (halo-reuse kernel) and impl = 4 (halo-reuse
kernel on CTA pairs, tcgen05 cta_group::2), timing each with CUDA events and comparing their outputs.  Measurement tool only; not part of the product path.
"""
import argparse
import ctypes as C
import json
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
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
    prog = eng.program(B, S, S)
    x = torch.randn(B, 3, S, S, device=dev)
    t = torch.full((B,), 500.0, device=dev)
    prog.run(x, t)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    convs = [(meta, a) for (name, meta), a in zip([oi for oi in prog.op_info if oi[0] == "conv"],
                                                   [k for k in prog.keep if isinstance(k, ConvArgs)])]
    tot = {2: 0.0, 3: 0.0, 4: 0.0}
    for meta, a in convs:
        oh, ow = {0: (a.h, a.w), 1: (a.h // 2, a.w // 2), 2: (a.h * 2, a.w * 2), 3: (a.h, a.w)}[a.mode]
        numel = a.n * oh * ow * a.cout
        res = {}
        outs = {}
        a_stats = a.out_stats
        a.out_stats = None   # timing replays would pile onto the real statistics
        for impl in (2, 3, 4):
            a.impl = impl
            rc = lib.dsg_conv(C.byref(a), st)
            if rc != 0:
                res[impl] = None
                continue
            torch.cuda.synchronize()
            # read the output through a tensor aliasing the raw pointer
            outs[impl] = _alias(a.out, numel, dev).clone()
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
            res[impl] = ms[len(ms) // 2]
        a.impl = 0
        a.out_stats = a_stats
        diff = None
        if 2 in outs and 3 in outs:
            diff = (outs[2].float() - outs[3].float()).abs().max().item()
            scale = outs[2].float().abs().max().item()
        row = {"mode": a.mode, "hw": [a.h, a.w], "cin": a.cin, "csc": a.csc1 + a.csc2, "cout": a.cout,
               "res": bool(a.residual), "flops": meta["flops"], "ms_stream": res[2], "ms_halo": res.get(3),
               "ms_pair": res.get(4), "pair_equal": bool(3 in outs and 4 in outs and torch.equal(outs[3], outs[4])),
               "maxdiff": diff, "scale": scale if diff is not None else None}
        rows.append(row)
        f = lambda m: "   n/a" if m is None else f"{m * 1000:7.1f}us {meta['flops'] / m / 1e9:7.1f}TF"
        print(f"mode{a.mode} {a.h:3d}x{a.w:<3d} cin{a.cin:5d}+{a.csc1 + a.csc2:<4d} cout{a.cout:4d} res{int(bool(a.residual))} "
              f"stream {f(res[2])}  halo {f(res.get(3))}  pair {f(res.get(4))}  maxdiff {diff} "
              f"pair==halo {row['pair_equal']}", flush=True)
        tot[2] += res[2]
        tot[3] += res[3] if res.get(3) is not None else res[2]
        tot[4] += res[4] if res.get(4) is not None else (res[3] if res.get(3) is not None else res[2])
    print(f"total: stream {tot[2]:.3f} ms, halo-where-available {tot[3]:.3f} ms, pair-where-available {tot[4]:.3f} ms")
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"batch": B, "size": S, "rows": rows, "total_stream_ms": tot[2], "total_halo_ms": tot[3],
                   "total_pair_ms": tot[4]},
                  open(args.out, "w"), indent=1)


def _alias(ptr, numel, dev):
    """fp16 tensor aliasing a raw device pointer (no ownership)."""
    class _Arr:
        __cuda_array_interface__ = {"shape": (numel,), "typestr": "<f2", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_Arr(), device=dev)


if __name__ == "__main__":
    main()
