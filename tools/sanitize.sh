#!/bin/bash
# compute-sanitizer (memcheck + racecheck + synccheck) over the smoke path and one small pipeline call.
# Usage (GPU box): bash tools/sanitize.sh [outdir]   -> <outdir>/sanitize_<tool>.log + sanitize_summary.txt
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
: > "$OUT/sanitize_summary.txt"
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_target.py > "$OUT/sanitize_$tool.log" 2>&1
  rc=$?
  echo "== $tool rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_$tool.log" | tail -1)" >> "$OUT/sanitize_summary.txt"
done
cat "$OUT/sanitize_summary.txt"
