#!/usr/bin/env python
"""Per-kernel roofline table from an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv`,
or a share-of-step table from a `--metrics gpu__time_duration.sum` launch list (`--csv --log-file`).

  python tools/ncu_table.py profiles/sweep_r1_final_raw.csv            # one row per captured launch
  python tools/ncu_table.py --launches profiles/launches_r1_final.csv  # time per kernel name

Peaks come from MEASURED_PEAKS.json when present (hbm GB/s), else 6570 GB/s (this pool's measured copy bandwidth)."""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbps", "hbm_GBps", "hbm_copy_gbps", "hbm"):
                if k in d:
                    v = d[k]
                    return float(v["burst"] if isinstance(v, dict) and "burst" in v else v)
        except Exception:
            pass
    return 6570.0


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def raw_table(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale=None):
        if name not in col:
            return float("nan")
        v = num(r[col[name]])
        u = units[col[name]].lower()
        if scale == "bytes":
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        if scale == "sec":
            v *= {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "s": 1, "second": 1}.get(u, 1)
        return v

    peak = hbm_peak()
    print(f"| kernel | grid | regs | ms | DRAM rd+wr MB | GB/s | % of {peak:.0f} GB/s | SM % | top pipe % |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\((?!int\)).*$", "", r[col["Kernel Name"]]).strip()
        t = get(r, "gpu__time_duration.sum", "sec")
        by = get(r, "dram__bytes_read.sum", "bytes") + get(r, "dram__bytes_write.sum", "bytes")
        pipes = [get(r, k) for k in col if k.startswith("sm__inst_executed_pipe_") and k.endswith("pct_of_peak_sustained_active")]
        pipes = [p for p in pipes if p == p]
        print(f"| {name} | {r[col['Grid Size']]} | {get(r, 'launch__registers_per_thread'):.0f} | {t * 1e3:.3f} | {by / 1e6:.1f} | "
              f"{by / t / 1e9:.0f} | {100 * by / t / 1e9 / peak:.1f} | {get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{max(pipes) if pipes else float('nan'):.0f} |")


def launch_table(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = num(r[vi]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
        k = re.sub(r"\((?!int\)).*$", "", r[ki]).strip()
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |")
    print(f"| total | {sum(v[0] for v in agg.values())} | {tot:.3f} | |")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--launches", action="store_true")
    a = ap.parse_args()
    (launch_table if a.launches else raw_table)(a.csv)
