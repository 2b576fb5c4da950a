#!/usr/bin/env python
"""Summarises an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>)
into the per-kernel table committed under profiles/.

    python tools/launch_list_summary.py gpurun_out/launches.csv "python bench.py --steps 3 --warmup 3 --no-cpu --no-workloads" > profiles/rN_launches_default.md
"""
import csv
import sys
from collections import defaultdict


def main(path, command):
    lines = [ln for ln in open(path, newline="") if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        a = agg[r["Kernel Name"]]
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list of `{command}`\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv`; times under ncu are serialised and cold-cache: compare shares, not absolutes.\n")
    print("| kernel | launches | mean us | total ms | share of GPU time |\n|---|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:110]}` | {n} | {us / n:.1f} | {us / 1e3:.2f} | {us / total:.3f} |")
    print(f"\n{sum(a[0] for a in agg.values())} launches, {total / 1e3:.1f} ms of GPU time in total.")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
