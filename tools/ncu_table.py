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
            for k in ("hbm_gbs", "hbm_gbps", "hbm_GBps", "hbm_copy_gbps", "hbm"):
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


def counters_table(path, frames=1.0):
    """Per-kernel table from a `--metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_*` pass over a frame
    (`--csv --log-file`): time, warp instructions, DRAM traffic, share of the frame's instruction budget."""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r[ix["ID"]], {"name": re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").strip()})
        d[r[ix["Metric Name"]]] = num(r[ix["Metric Value"]])
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for d in per.values():
        a = agg[re.sub(r"<.*", "", d["name"])]
        a["n"] += 1
        a["ms"] += d.get("gpu__time_duration.sum", 0) / 1e6
        a["inst"] += d.get("smsp__inst_executed.sum", 0)
        a["dram"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        a["issue_w"] += d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) * d.get("gpu__time_duration.sum", 0)
    tot_ms = sum(a["ms"] for a in agg.values()); tot_i = sum(a["inst"] for a in agg.values())
    peak = hbm_peak()
    print(f"{len(per)} launches = {frames:.2f} frame(s): {tot_ms / frames:.1f} ms and {tot_i / frames / 1e9:.2f} G warp instructions per frame alone; "
          f"at 148 SMs x 4 issue slots x 1.965 GHz = 1163 G/s the instructions alone need {tot_i / frames / 1163e9 * 1e3:.1f} ms per frame\n")
    print("| kernel | launches | ms | G warp-inst | % of inst | issue active % | DRAM GB | GB/s | % of HBM peak |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"]):
        print(f"| {k} | {int(a['n'])} | {a['ms']:.2f} | {a['inst'] / 1e9:.3f} | {100 * a['inst'] / tot_i:.1f} | {a['issue_w'] / max(a['ms'] * 1e6, 1):.0f} | "
              f"{a['dram'] / 1e9:.2f} | {a['dram'] / 1e9 / max(a['ms'] / 1e3, 1e-9):.0f} | {100 * a['dram'] / 1e9 / max(a['ms'] / 1e3, 1e-9) / peak:.0f} |")


def traffic_json(path, kernel_regex, out):
    """dram bytes per launch of the launches whose name matches, from a raw `--set full` export -> JSON for bench.py."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    vals = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not re.search(kernel_regex, r[col["Kernel Name"]]):
            continue
        b = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += num(r[col[name]]) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[col[name]].lower(), 1)
        vals.append(b)
    j = {"kernel_regex": kernel_regex, "source": os.path.relpath(path, ROOT), "launches": len(vals), "dram_bytes_per_launch": sum(vals) / max(len(vals), 1),
         "note": f"mean dram__bytes_read.sum + dram__bytes_write.sum of {len(vals)} captured launch(es) ({os.path.basename(path)}, ncu --set full)"}
    json.dump(j, open(out, "w"), indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--launches", action="store_true")
    ap.add_argument("--counters", action="store_true", help="per-kernel time / instructions / DRAM table from a --metrics pass")
    ap.add_argument("--frames", type=float, default=1.0, help="frames covered by the --counters capture")
    ap.add_argument("--traffic", metavar="KERNEL_REGEX", help="write the dominant kernel's DRAM bytes per launch as JSON (--out)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_dominant_kernel_traffic.json"))
    a = ap.parse_args()
    if a.traffic:
        traffic_json(a.csv, a.traffic, a.out)
    elif a.counters:
        counters_table(a.csv, a.frames)
    else:
        (launch_table if a.launches else raw_table)(a.csv)
