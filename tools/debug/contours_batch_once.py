#!/usr/bin/env python
"""Batched device contours (256 x 1080p design shape masks, device-resident) kernel by kernel, with the border segments
off / on for batches (llfe_set_option("contour_segments", 1 | 2)).

    python tools/debug/contours_batch_once.py [--batch 256]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import low_level_feature_extraction_b200 as pkg
    from low_level_feature_extraction_b200.synth import design_image

    eng = pkg.engine(0)
    imgs = np.stack([design_image(1080, 1920, s) for s in range(32)])
    batch = torch.from_numpy(imgs).cuda().repeat(a.batch // 32, 1, 1, 1)
    masks = eng.shape_mask(batch)
    out = {}
    ref = None
    for mode in (1, 2, 1, 2):
        eng.ctx.set_option("contour_segments", mode)
        hdr, pts, cnt = eng.contours_external(masks, 200, 1024, 8192)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            eng.ctx.call("llfe_contours_external", masks, a.batch, 1080, 1920, 200, hdr, 1024, pts, 8192, cnt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        eng.ctx.profile_begin()
        eng.ctx.call("llfe_contours_external", masks, a.batch, 1080, 1920, 200, hdr, 1024, pts, 8192, cnt)
        torch.cuda.synchronize()
        prof = eng.ctx.profile_end()
        # same contours either way (slot order differs between runs: compare per image as sets of (start, npts, area))
        h = hdr.cpu().numpy()
        c = cnt.cpu().numpy()
        sig = [sorted(map(tuple, h[i, :min(int(c[i, 0]), 1024)][:, [0, 1, 3, 4, 5, 6, 8, 9]].tolist())) for i in range(a.batch)]
        if ref is None:
            ref = sig
        out.setdefault(f"contour_segments={mode}", []).append(
            {"ms_per_call": round(ms, 3), "same_as_first": sig == ref, "kernels_ms": {k: round(v["ms"], 4) for k, v in prof.items()}})
    eng.ctx.set_option("contour_segments", 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
