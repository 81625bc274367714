"""Summarise an ncu report: per-source-line instruction / stall-sample shares, opcode mix, stall reasons.
usage: python scratch/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for i, h in enumerate(hdr):
    if h in want: print(f"{h:70s} {rows[1][i]:>14s} {vals[i]}")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ix, sx = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = collections.OrderedDict(); ops = collections.Counter(); stalls = collections.Counter(); tot = 0; tots = 0
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    if r[0].strip():   # source line row
        cur = (r[0], r[1]); lines.setdefault(cur, [0, 0]); continue
    try: n = int(r[ix]); s = int(r[sx])
    except ValueError: continue
    if cur: lines[cur][0] += n; lines[cur][1] += s
    tot += n; tots += s
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3]); op = m.group(2).split(".")[0] if m else "?"
    if op == "IMAD" and "MOV" in r[3]: op = "IMAD.MOV"
    ops[op] += n
    for i in stall_cols:
        try: stalls[hdr[i]] += int(r[i])
        except ValueError: pass
print(f"\ntotal warp instructions {tot}, samples {tots}")
print("\n-- top source lines by instructions executed")
for (ln, src), (n, s) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{ln:>5s} {100*n/tot:5.1f}% inst {100*s/max(tots,1):5.1f}% smp  {src.strip()[:105]}")
print("\n-- opcode mix"); print("  ".join(f"{k}:{100*v/tot:.1f}%" for k, v in ops.most_common(22)))
ss = sum(stalls.values()) or 1
print("\n-- stall reasons"); print("  ".join(f"{k[6:]}:{100*v/ss:.1f}%" for k, v in stalls.most_common(9)))
