"""Prototype of the segmented border follower (k_contours.cu: k_ct_segments + the hopping leader), in plain Python,
against cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE).  States are (pixel, move direction into the pixel); a HEAD
is a state on a cut row (y % C == 0) entered by a move with a vertical component.  Every candidate head is followed to
the next head (a segment); the leader walks from the contour's start state and hops from head to head.

    python tools/debug/contour_segments_proto.py
"""
import sys

import cv2
import numpy as np

DX = [1, 1, 0, -1, -1, -1, 0, 1]
DY = [0, -1, -1, -1, 0, 1, 1, 1]


def nbits(m, x, y):
    nb = 0
    for k in range(8):
        if m[y + DY[k], x + DX[k]]:
            nb |= 1 << k
    return nb


def step_dir(nb, din):
    s = (din + 4) & 7
    for j in range(1, 9):
        if (nb >> ((s + j) & 7)) & 1:
            return (s + j) & 7
    raise AssertionError


def is_head(x, y, d, C):
    """padded (x, y) entered by a move in direction d: on a cut row by a vertical move or on a cut column by a horizontal one"""
    return (DY[d] != 0 and (y - 1) % C == 0) or (DX[d] != 0 and (x - 1) % C == 0)


def segments(m, C, L=1 << 30):
    """m: padded 0/1 mask.  -> {(x, y, d): (end state or None, [vertices])} for every candidate head"""
    h, w = m.shape
    tab = {}
    for y in range(1, h - 1):
        for x in range(1, w - 1):
            if not m[y, x]:
                continue
            for d in range(8):
                if not is_head(x, y, d, C) or not m[y - DY[d], x - DX[d]]:
                    continue
                cx, cy, din, pts, end = x, y, d, [], None
                for _ in range(L):
                    sn = step_dir(nbits(m, cx, cy), din)
                    if sn != din:
                        pts.append((cx - 1, cy - 1))
                    cx, cy, din = cx + DX[sn], cy + DY[sn], sn
                    if is_head(cx, cy, din, C):
                        end = (cx, cy, din)
                        break
                tab[(x, y, d)] = (end, pts)
    return tab


def leader(m, x0, y0, tab, C, stats):
    """x0, y0: padded coordinates of the raster-first pixel of a component"""
    nb = nbits(m, x0, y0)
    if nb == 0:
        return [(x0 - 1, y0 - 1)]
    s = 3
    while not (nb >> s) & 1:
        s = (s - 1) & 7
    x1, y1 = x0 + DX[s], y0 + DY[s]
    X, Y, din = x0, y0, s ^ 4
    pts, first_head, final = [], None, False
    while True:
        if not final and is_head(X, Y, din, C):
            hid = (X, Y, din)
            if first_head is None:
                first_head = hid
            while True:
                end, seg = tab[hid]
                if end is None:
                    break
                if end == first_head:
                    final = True
                    break
                pts += seg
                stats[0] += 1
                hid = end
            X, Y, din = hid
        sn = step_dir(nbits(m, X, Y), din)
        stats[1] += 1
        if sn != din:
            pts.append((X - 1, Y - 1))
        X4, Y4 = X + DX[sn], Y + DY[sn]
        if (X4, Y4) == (x0, y0) and (X, Y) == (x1, y1):
            break
        X, Y, din = X4, Y4, sn
    return pts


def find_external_segmented(mask, C, L=1 << 30, stats=None):
    sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
    from oracle import contours as oc

    stats = stats if stats is not None else [0, 0]
    m = np.pad((mask != 0).astype(np.uint8), 1)
    tab = segments(m, C, L)
    out = []
    for (x, y) in oc.external_starts_ideal(mask):
        out.append(leader(m, x + 1, y + 1, tab, C, stats))
    return out


def main():
    rng = np.random.default_rng(0)
    bad = 0
    for trial in range(300):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        dens = rng.choice([0.1, 0.3, 0.5, 0.7, 0.9])
        mask = (rng.random((h, w)) < dens).astype(np.uint8) * 255
        if trial % 5 == 0:      # thick blobs
            mask = cv2.dilate(mask, np.ones((3, 3), np.uint8))
        C = int(rng.choice([1, 2, 4, 8]))
        L = int(rng.choice([3, 10, 1 << 30]))
        stats = [0, 0]
        got = find_external_segmented(mask, C, L, stats)
        ref, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        a = sorted(tuple(map(tuple, c.reshape(-1, 2).tolist())) for c in ref)
        b = sorted(tuple(c) for c in got)
        if a != b:
            bad += 1
            print("MISMATCH", trial, h, w, dens, C, L)
    print("bad", bad)


if __name__ == "__main__":
    main()
