#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report (one or more kernels) as markdown:
headline metrics per kernel, hottest source lines, opcode mix, stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] [--out profiles/x.md] [--traffic profiles/ncu_traffic.json]

Runs here (no GPU needed): it only reads the report with `ncu -i`.
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
UNIT_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def short(name: str) -> str:
    m = re.search(r"(k_[A-Za-z0-9_]+)", name)
    if not m:
        return name[:40]
    n = m.group(1)
    if n == "k_kmeans_fast" and re.search(r"k_kmeans_fast<\d+, (0|false|\(bool\)0)>", name):
        n += "_global"   # second instantiation: lists that do not fit in shared memory
    return n


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr) if i < len(r)}
        out.append(d)
    return out


def source_pages(rep):
    """-> one page per captured launch: {"name", "hdr", "rows"} (blocks of all source files merged)."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur, path = [], None, None
    for r in rows:
        if r and r[0] == "File Path":
            path = r[1] if len(r) > 1 else ""
            cur = None
        elif r and r[0] in ("Function Name", "Kernel Name"):
            cur = {"name": r[1], "path": path, "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                if r and r[0] in ("Line No", "Address", "#"):
                    cur["hdr"] = r
            else:
                cur["rows"].append([path] + r if r and r[0].strip() else r)
    pages = []
    for b in blocks:
        if pages and pages[-1]["name"] == b["name"] and b["path"] not in pages[-1]["paths"]:
            pages[-1]["rows"] += b["rows"]
            pages[-1]["paths"].add(b["path"])
        else:
            pages.append({"name": b["name"], "hdr": b["hdr"], "rows": list(b["rows"]), "paths": {b["path"]}})
    return pages


def summarise_source(page, top):
    hdr = page["hdr"]
    if not hdr or "Instructions Executed" not in hdr:
        return ["(no source page)"]
    ix, sx = hdr.index("Instructions Executed"), hdr.index("# Samples")
    src_col = hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and "not_issued" not in h.lower()]
    lines = collections.OrderedDict()
    ops, stalls = collections.Counter(), collections.Counter()
    tot = tots = 0
    cur = None
    is_line_mode = hdr[0] == "Line No"
    for r in page["rows"]:
        if is_line_mode and len(r) == len(hdr) + 1 and r[1].strip():   # a source-line row (file path prepended)
            cur = (os.path.basename(r[0]) + ":" + r[1], r[2])
            lines.setdefault(cur, [0, 0])
            continue
        if len(r) < len(hdr):
            continue
        try:
            n, s = int(r[ix]), int(r[sx])
        except ValueError:
            continue
        if cur:
            lines[cur][0] += n
            lines[cur][1] += s
        tot += n
        tots += s
        sass = r[src_col] if not is_line_mode else r[3]
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
        op = m.group(2) if m else "?"
        base = op.split(".")[0]
        if base == "IMAD" and "MOV" in op:
            base = "IMAD.MOV"
        ops[base] += n
        for i in stall_cols:
            try:
                stalls[hdr[i]] += int(r[i])
            except ValueError:
                pass
    out = [f"warp instructions executed: {tot:,}; stall samples: {tots:,}", ""]
    if lines:
        out += ["| line | % inst | % samples | source |", "|---|---|---|---|"]
        for (ln, src), (n, s) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            out.append(f"| {ln} | {100 * n / max(tot, 1):.1f} | {100 * s / max(tots, 1):.1f} | `{src.strip()[:110].replace('|', '¦')}` |")
        out.append("")
    out.append("opcode mix: " + "  ".join(f"{k} {100 * v / max(tot, 1):.1f}%" for k, v in ops.most_common(24)))
    ss = sum(stalls.values()) or 1
    out.append("")
    out.append("stall reasons: " + "  ".join(f"{k[6:]} {100 * v / ss:.1f}%" for k, v in stalls.most_common(10)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--out")
    ap.add_argument("--traffic", help="JSON file to update with dram bytes per launch per kernel")
    ap.add_argument("--title", default=None)
    ap.add_argument("--images", type=int, default=None, help="images per launch in the captured run (stored with the traffic)")
    ap.add_argument("--fresh", action="store_true", help="rewrite the traffic file instead of merging into it")
    a = ap.parse_args()
    raws = raw_page(a.rep)
    pages = source_pages(a.rep)
    md = [f"# {a.title or os.path.basename(a.rep)}", "",
          f"Source report: `{a.rep}` (`ncu --set full --clock-control none --import-source on`; per-launch figures are "
          "cold-cache and serialised -- compare shares and ratios, not absolute times).", ""]
    traffic = {}
    seen = collections.Counter()
    for i, d in enumerate(raws):
        name = short(d.get("Kernel Name", ("?", ""))[0])
        seen[name] += 1
        md += [f"## {name} (launch #{seen[name]} in the capture)", "", "| metric | value | unit |", "|---|---|---|"]
        for m in METRICS:
            if m in d:
                md.append(f"| {m} | {d[m][0]} | {d[m][1]} |")
        try:
            rd = float(d["dram__bytes_read.sum"][0]) * UNIT_BYTES.get(d["dram__bytes_read.sum"][1], 1)
            wr = float(d["dram__bytes_write.sum"][0]) * UNIT_BYTES.get(d["dram__bytes_write.sum"][1], 1)
            ms = float(d["gpu__time_duration.sum"][0]) * UNIT_MS.get(d["gpu__time_duration.sum"][1], 1)
            md.append(f"| dram traffic per launch | {(rd + wr) / 1e6:.3f} | MB |")
            md.append(f"| dram GB/s in this (serialised) launch | {(rd + wr) / 1e9 / (ms / 1e3):.1f} | GB/s |")
            traffic.setdefault(name, []).append((rd + wr, ms))
        except (KeyError, ValueError):
            pass
        md.append("")
        if i < len(pages):
            md += summarise_source(pages[i], a.top)
            md.append("")
    # average per kernel, ignoring launches that exit at once (tiers that own no image in the capture)
    for name, recs in list(traffic.items()):
        longest = max(ms for _, ms in recs)
        keep = [(b, ms) for b, ms in recs if ms >= 0.05 * longest]
        traffic[name] = {"bytes_per_launch": sum(b for b, _ in keep) / len(keep), "ms": sum(ms for _, ms in keep) / len(keep),
                         "n": len(keep)}
    text = "\n".join(md) + "\n"
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    else:
        print(text)
    if a.traffic:
        old = {}
        if os.path.exists(a.traffic) and not a.fresh:
            with open(a.traffic) as f:
                old = json.load(f)
        for k, v in traffic.items():
            old[k] = {"bytes_per_launch": v["bytes_per_launch"], "ncu_ms_per_launch": v["ms"], "launches_averaged": v["n"],
                      "report": os.path.basename(a.rep), "images_per_launch": a.images}
        with open(a.traffic, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
