#!/usr/bin/env python
"""BASELINE.json configs[4]: GroupNorm+SiLU / attention / conv micro-benchmark sweep, 64 -> 512 px, against the measured
HBM and tensor-pipe rooflines (MEASURED_PEAKS.json).  One JSON document on stdout / --out.

    python tools/op_sweep.py --out gpurun_out/op_sweep.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from bench import peaks  # noqa: E402
from drivescenegen_b200 import ops  # noqa: E402


def timed(fn, reps=5, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)   # 256 MB > L2: evicts the previous iteration's data
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--sizes", default="64,128,256,512")
    ap.add_argument("--batches", default="1,8,16")
    ap.add_argument("--raster", action="store_true", help="also sweep the byte-image kernels (raster.cu)")
    ap.add_argument("--raster-only", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pk = peaks()
    flush = torch.zeros(64 << 20, device=dev)
    rows = []
    chans = (64, 128, 256, 512)
    if args.raster_only:
        args.raster = True
    for size in ([] if args.raster_only else [int(s) for s in args.sizes.split(",")]):
        for b in [int(s) for s in args.batches.split(",")]:
            if size == 512 and b > 8:
                continue
            for lvl, c in enumerate(chans):
                h = size >> lvl
                x = torch.randn(b, h, h, c, device=dev).half()
                gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
                st = ops.gn_stats(x)
                ms = timed(lambda: ops.group_norm(x, None, gamma, beta, 32, 1e-5, 1, stats1=st), flush=flush)
                nbytes = x.numel() * 2 * 2
                rows.append({"op": "groupnorm_silu_apply", "px": size, "batch": b, "hw": h, "c": c, "ms": ms,
                             "GBs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / pk["hbm_gbs"]})
                # 3x3 conv c -> c at this level (the "4.832 GFLOP" family of SURVEY App. C)
                w = torch.randn(c, c, 3, 3, device=dev) / (3.0 * c ** 0.5)
                wp = ops.pack_conv_weight(0, w)
                ms = timed(lambda: ops.conv(0, x, wp, c), flush=flush)
                fl = 2.0 * b * h * h * c * c * 9
                rows.append({"op": "conv3x3", "px": size, "batch": b, "hw": h, "cin": c, "cout": c, "ms": ms,
                             "TFLOPs": fl / ms / 1e9, "frac_tensor": fl / ms / 1e9 / pk["tflops_burst"]})
            tokens = (size // 8) ** 2
            if tokens >= 64:
                qkv = torch.randn(b, tokens, 3 * 512, device=dev).half()
                ms = timed(lambda: ops.attention(qkv, 64, 8), flush=flush)
                rows.append({"op": "attention_64h_d8", "px": size, "batch": b, "tokens": tokens, "ms": ms,
                             "Gexp_s": b * 64 * tokens * tokens / ms / 1e6,
                             "TFLOPs": 4.0 * b * tokens * tokens * 512 / ms / 1e9})
    if args.raster:
        from drivescenegen_b200.hostapi import raster
        import numpy as np
        rng = np.random.default_rng(0)
        for size in [int(s) for s in args.sizes.split(",")]:
            for b in ((16, 256, 1024) if size <= 256 else (16, 64, 256)):
                img = rng.integers(0, 256, (b, size, size, 3), dtype=np.uint8)
                bg = rng.random((b, size, size)) < 0.9          # BEV rasters: one dominant background value
                img[..., 0] = np.where(bg, 127, img[..., 0])
                img[..., 1] = np.where(bg, 128, img[..., 1])
                img = torch.from_numpy(img).to(dev)
                px = b * size * size
                out = torch.empty((b, 3, size, size), device=dev)
                ms = timed(lambda: raster.image_to_sample(img, out=out), flush=flush)
                nb = px * (3 + 12)
                rows.append({"op": "image_to_sample", "px": size, "batch": b, "ms": ms, "GBs": nb / ms / 1e6,
                             "frac_hbm": nb / ms / 1e6 / pk["hbm_gbs"]})
                ms = timed(lambda: raster.gray_masks(img), flush=flush)
                nb = px * (3 + 3 + 1)     # histogram pass + mask pass read the image, mask pass writes 1 B/px
                rows.append({"op": "gray_mask(hist+mask)", "px": size, "batch": b, "ms": ms, "GBs": nb / ms / 1e6,
                             "frac_hbm": nb / ms / 1e6 / pk["hbm_gbs"]})
                ms = timed(lambda: raster.agent_threshold(out), flush=flush)
                nb = px * (4 + 1)
                rows.append({"op": "agent_threshold", "px": size, "batch": b, "ms": ms, "GBs": nb / ms / 1e6,
                             "frac_hbm": nb / ms / 1e6 / pk["hbm_gbs"]})
    doc = {"peaks": pk, "note": "min of 5 CUDA-event timings, L2 flushed between iterations; conv fraction is of the "
                                "measured BURST bf16 peak (kernel timed alone)", "rows": rows}
    txt = json.dumps(doc, indent=1)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        open(args.out, "w").write(txt)
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
