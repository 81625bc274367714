#!/usr/bin/env python
"""Kernel timeline of ONE pipeline step (256 x 1080p, device-resident) from CUPTI (torch.profiler): which kernels of the
two chains of `llfe_analyze` -- masks on one stream, colours on the other -- actually ran at the same time.

    python tools/timeline.py [images] > profiles/timeline_r2.json

Writes {"kernels": [[name, stream, start_us, dur_us], ...], "summary": {...}}; the summary has the step's span, the busy
time of each stream and the time during which both streams had a kernel running."""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image  # noqa: E402


def union(iv):
    iv = sorted(iv)
    tot, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    return tot + (cur_b - cur_a if cur_b is not None else 0.0)


def overlap(a, b):
    """time covered by both interval sets"""
    return union(a) + union(b) - union(a + b)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    distinct = min(n, 32)
    base = np.stack([design_image(1080, 1920, s) for s in range(distinct)])
    bgr = torch.from_numpy(np.concatenate([base] * (n // distinct))).cuda()
    ba = BatchAnalyzer(0, 1080, 1920, BatchConfig())
    out = ba.alloc_outputs(n)
    for _ in range(3):
        ba.run_device(bgr, out, resolve=False)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ba.run_device(bgr, out, resolve=False)
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.name.find("k_") >= 0:
            ev.append((e.name, e.time_range.start, e.time_range.end))
    # torch's FunctionEvent does not expose the stream: take it from the chrome trace
    path = "/tmp/llfe_trace.json"
    prof.export_chrome_trace(path)
    tr = json.load(open(path))
    ks = [(x["name"], int(x["args"].get("stream", -1)), float(x["ts"]), float(x["dur"])) for x in tr["traceEvents"]
          if x.get("cat") == "kernel"]
    t0 = min(k[2] for k in ks)
    def short(name):
        m = re.search(r"(k_[A-Za-z0-9_]+)", name)
        return m.group(1) if m else name[:40]

    ks = sorted([(short(k[0]), k[1], round(k[2] - t0, 2), round(k[3], 2)) for k in ks], key=lambda k: k[2])
    streams = sorted({k[1] for k in ks})
    per = {s: [(k[2], k[2] + k[3]) for k in ks if k[1] == s] for s in streams}
    span = max(k[2] + k[3] for k in ks)
    summary = {"images": n, "span_us": round(span, 1), "kernels": len(ks),
               "busy_us_per_stream": {str(s): round(union(per[s]), 1) for s in streams},
               "sum_of_kernel_durations_us": round(sum(k[3] for k in ks), 1)}
    if len(streams) >= 2:
        a, b = sorted(streams, key=lambda s: -union(per[s]))[:2]
        summary["both_streams_busy_us"] = round(overlap(per[a], per[b]), 1)
        summary["any_stream_busy_us"] = round(union(per[a] + per[b]), 1)
    by = {}
    for k in ks:
        d = by.setdefault(k[0], {"launches": 0, "us": 0.0, "stream": k[1]})
        d["launches"] += 1
        d["us"] = round(d["us"] + k[3], 1)
    summary["per_kernel"] = by
    print(json.dumps({"summary": summary, "kernels": ks}))


if __name__ == "__main__":
    main()
