"""The contour oracle (oracle/contours.py) against the installed cv2 binary (CPU only)."""
import cv2
import numpy as np
import pytest

from oracle import contours as oc


def random_mask(rng, h, w):
    dens = rng.choice([0.1, 0.25, 0.4, 0.5, 0.6, 0.75, 0.9, 0.97])
    m = (rng.random((h, w)) < dens).astype(np.uint8) * 255
    kind = rng.integers(0, 4)
    if kind == 1:
        m = cv2.dilate(m, np.ones((3, 3), np.uint8))
    elif kind == 2:
        m = cv2.erode(m, np.ones((3, 3), np.uint8))
    elif kind == 3:
        m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, np.ones((3, 3), np.uint8))
    return m


def cv_contours(m):
    ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return [[(int(p[0][0]), int(p[0][1])) for p in c] for c in ref]


def test_scan_restatement_equals_cv2_point_for_point():
    rng = np.random.default_rng(0)
    for _ in range(1500):
        h, w = rng.integers(1, 28, 2)
        m = random_mask(rng, int(h), int(w))
        assert oc.find_external(m) == cv_contours(m), m // 255


def test_closed_form_of_the_external_set_equals_cv2():
    rng = np.random.default_rng(1)
    for _ in range(4000):
        h, w = rng.integers(1, 40, 2)
        m = random_mask(rng, int(h), int(w))
        ref = cv_contours(m)
        starts = sorted((c[0] for c in ref), key=lambda p: (p[1], p[0]))
        assert oc.external_starts_ideal(m) == starts, m // 255


def test_area_formula_equals_cv2():
    rng = np.random.default_rng(2)
    for _ in range(200):
        m = random_mask(rng, 30, 40)
        ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for c in ref:
            pts = [(int(p[0][0]), int(p[0][1])) for p in c]
            assert abs(oc.contour_area2(pts)) == int(round(2 * cv2.contourArea(c)))


def test_segmented_follower_prototype_equals_cv2():
    """the algorithm behind k_ct_segments + the hopping leader (tools/debug/contour_segments_proto.py): heads on cut rows /
    columns, every candidate followed to the next head, the leader hops; segments capped at a few steps are left to the
    leader.  Point for point against cv2."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "debug"))
    import contour_segments_proto as proto

    rng = np.random.default_rng(5)
    hops = 0
    for trial in range(60):
        h, w = int(rng.integers(1, 36)), int(rng.integers(1, 36))
        m = random_mask(rng, h, w)
        stats = [0, 0]
        got = proto.find_external_segmented(m, int(rng.choice([1, 2, 4, 8])), int(rng.choice([3, 10, 1 << 30])), stats)
        hops += stats[0]
        assert sorted(map(tuple, got)) == sorted(map(tuple, cv_contours(m))), (trial, h, w)
    assert hops > 500     # the hop path is what ran
