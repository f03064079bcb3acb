#!/usr/bin/env python
"""Per-launch summary of an `ncu --page raw --csv` export: duration, DRAM bytes, DRAM %, tensor-pipe %.

    python tools/ncu_raw_summary.py gpurun_out/prof_conv_raw.csv [--json out.json]
"""
import csv
import json
import re
import sys

COLS = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    ki = hdr.index("Kernel Name")
    out = []
    for r in data:
        name = re.sub(r"\(.*", "", r[ki])[:60]
        rec = {"kernel": name}
        for k, i in idx.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            rec[k] = v * UNIT.get(units[i], 1.0)
        out.append(rec)
    print(f"{'#':>3s} {'kernel':60s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'dram%':>6s} {'tensor%':>7s} {'GB/s':>7s}")
    for i, r in enumerate(out):
        tr = r.get("dram_rd", 0) + r.get("dram_wr", 0)
        print(f"{i:3d} {r['kernel']:60s} {r.get('dur_us', 0):8.1f} {r.get('dram_rd', 0) / 1e6:8.1f} "
              f"{r.get('dram_wr', 0) / 1e6:8.1f} {r.get('dram_pct', 0):6.1f} {r.get('tensor_pct', 0):7.1f} "
              f"{tr / max(r.get('dur_us', 1), 1e-9) / 1e3:7.0f}")
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
