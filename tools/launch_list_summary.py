#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) per kernel and put
the shares next to the CUDA-event shares of a bench line.

    python tools/launch_list_summary.py profiles/launches_r1_v31.csv profiles/bench_r1_v31.json > profiles/launches_r1_v31.md
"""
import csv
import json
import re
import sys
from collections import OrderedDict


def main():
    path, bench = sys.argv[1], sys.argv[2]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", re.sub(r"^.*::", "", r["Kernel Name"].split("(")[0]))
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((name, ms))
    agg = OrderedDict()
    for name, ms in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    tot = sum(v[1] for v in ours.values())
    b = json.loads(open(bench).read().strip().splitlines()[-1])
    bk = {k: v["ms_per_step"] for k, v in b.get("kernels", {}).items()}
    btot = sum(bk.values()) or 1.0
    print(f"# ncu launch list next to `{bench}`\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` on `python bench.py --steps 2 --warmup 3 "
          "--no-cpu-baseline --no-e2e`. Per-launch times under ncu are cold-cache and serialised: compare the SHARES with "
          "the shares the bench line (CUDA events in the timed region) reports for the same kernels.\n")
    print("| kernel | launches | total ms (ncu) | share (ncu) | share (bench, CUDA events) |\n|---|---|---|---|---|")
    for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {100 * bk.get(k, 0.0) / btot:.1f}% |")
    other = {k: v for k, v in agg.items() if not k.startswith("k_")}
    if other:
        print("\nOther launches in the capture (torch fills / copies): " +
              ", ".join(f"{k} x{n} ({ms:.3f} ms)" for k, (n, ms) in other.items()))


if __name__ == "__main__":
    main()
