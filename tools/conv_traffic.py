#!/usr/bin/env python
"""profiles/conv_traffic_b<B>_<S>.json — the DRAM traffic of the conv launches of one denoise step, from an
`ncu --set full` capture of THIS code (bench.py copies it into `roofline.traffic` and names the capture's commit).

    python tools/conv_traffic.py <raw.csv from `ncu --page raw --csv`> <batch> <size> <commit> <out.json> [tensor-weighted]
"""
import json
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import csv
import re

from ncu_raw_summary import COLS, UNIT


def main():
    path, batch, size, commit, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    ki = hdr.index("Kernel Name")
    recs = []
    for r in data:
        if "igemm" not in r[ki]:
            continue
        rec = {"kernel": re.sub(r"\(.*", "", r[ki])[:60]}
        for k, i in idx.items():
            try:
                rec[k] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            except ValueError:
                pass
        recs.append(rec)
    total = sum(r.get("dram_rd", 0) + r.get("dram_wr", 0) for r in recs)
    dur = sum(r.get("dur_us", 0) for r in recs)
    tensor_w = sum(r.get("tensor_pct", 0) * r.get("dur_us", 0) for r in recs) / max(dur, 1e-9)
    json.dump({"workload": f"{size}x{size}x3, batch {batch}, one U-Net forward", "launches": len(recs),
               "dram_bytes_total": total, "dram_bytes_per_launch": total / max(len(recs), 1),
               "tensor_pipe_pct_time_weighted": tensor_w, "duration_us_under_ncu": dur, "commit": commit,
               "source": f"ncu --set full --clock-control none of tools/profile_step.py at commit {commit}"},
              open(out, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main()
