"""LK stage (src/Frame.cc:445): the CPU oracle restatement of calcOpticalFlowPyrLK against cv2-generated golden vectors (and cv2 live).
Tolerance-based: OpenCV accumulates its float sums in SIMD order, the restatement in scalar order."""
import os

import numpy as np
import pytest

import oracle as O

# tolerances in pixels on the tracked position
TOL_MAX = 0.02     # every point (the iteration may stop one step earlier/later)
TOL_P99 = 2e-3
TOL_MEDIAN = 2e-4


def _check(mine, ref):
    d = np.abs(mine - ref).max(1)
    assert np.median(d) <= TOL_MEDIAN and np.quantile(d, 0.99) <= TOL_P99 and d.max() <= TOL_MAX, (np.median(d), np.quantile(d, 0.99), d.max())


def test_oracle_lk_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'lk_320x240.npz'))
    lv = g['cur']
    for l in (1, 2, 3):
        lv = O.lk_pyr_level(g['cur'], l)
        assert int(lv.astype(np.uint64).sum()) == int(g['pyr_sums'][l - 1])
    assert np.array_equal(lv, g['pyr3'])                       # cv::pyrDown is integer arithmetic: bit-exact
    _check(O.lk_track(g['cur'], g['prev'], g['pts']), g['tracked'])


def test_oracle_lk_against_cv2_live():
    cv2 = pytest.importorskip('cv2')
    from pysgs import synth
    frames, _ = synth.stream_s2(3, 640, 480, seed=2)
    for a, b in ((1, 0), (2, 1)):
        k, _ = O.extract(frames[a])
        pts = np.stack([k['x'], k['y']], 1).astype(np.float32)
        ref, _, _ = cv2.calcOpticalFlowPyrLK(frames[a], frames[b], pts, None, winSize=(21, 21), maxLevel=3,
                                             criteria=(cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01))
        _check(O.lk_track(frames[a], frames[b], pts), ref.reshape(-1, 2))
