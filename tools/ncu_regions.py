#!/usr/bin/env python
"""Shares of executed warp instructions and stall samples per source region of one kernel in an ncu report
(`--import-source on`):  python tools/ncu_regions.py REPORT KERNEL_REGEX FILE lo:hi:name [lo:hi:name ...]"""
import collections
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep, kern, fname = sys.argv[1:4]
    regions = [(int(a), int(b), c) for a, b, c in (r.split(":") for r in sys.argv[4:])]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{kern}"], capture_output=True, text=True).stdout
    data, path, name, hdr = collections.OrderedDict(), None, "", None
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1]
        elif r[0] in ("Function Name", "Kernel Name"):
            name = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0].strip().isdigit() and hdr and len(r) > 8:
            a = data.setdefault((name, path.split("/")[-1], int(r[0])), [0, 0, r[1]])
            a[0] += num(r[hdr.index("Instructions Executed")])
            a[1] += num(r[hdr.index("# Samples")])
    names = collections.Counter()
    for (n, f, l), v in data.items():
        names[n] += v[0]
    main_name = names.most_common(1)[0][0]
    print("kernel:", main_name[:120])
    d = {(f, l): v for (n, f, l), v in data.items() if n == main_name}
    tot = sum(v[0] for v in d.values())
    tots = sum(v[1] for v in d.values())
    agg, aggs = collections.Counter(), collections.Counter()
    for (f, ln), v in d.items():
        key = "other: " + f
        if f == fname:
            for lo, hi, n in regions:
                if lo <= ln < hi:
                    key = n
                    break
        agg[key] += v[0]
        aggs[key] += v[1]
    print(f"total warp instructions {tot}, stall samples {tots}")
    for k in agg:
        print(f"{k:34s} inst {100 * agg[k] / tot:5.1f} %   samples {100 * aggs[k] / tots:5.1f} %")


if __name__ == "__main__":
    main()
