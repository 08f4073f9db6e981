#!/usr/bin/env python3
"""Golden vectors for the LK stage (src/Frame.cc:445): the REAL cv2.calcOpticalFlowPyrLK with the reference's parameters
(winSize 21x21, maxLevel 3, COUNT|EPS 30 / 0.01) on seeded synthetic frame pairs, plus cv2.pyrDown levels.
Run in the build container (needs cv2):  python tests/golden/make_golden_lk.py"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, 'sg-slam_b200'), os.path.join(ROOT, 'oracle')]
from pysgs import synth  # noqa: E402
import oracle as O       # noqa: E402  (only to pick realistic keypoint positions; the vectors themselves come from cv2)


def main():
    frames, _ = synth.stream_s2(2, 320, 240, seed=31, tex_w=512, tex_h=384, person=False)
    cur, prev = frames[1], frames[0]
    k, _ = O.extract(cur, O.params(300, 1.2, 5, 20, 7))
    rng = np.random.RandomState(0)
    pts = np.stack([k['x'], k['y']], 1).astype(np.float32)
    # add a few hard cases: near the image border, on flat regions, sub-pixel positions
    extra = np.array([[1.5, 1.5], [318.2, 2.7], [3.3, 236.9], [160.25, 120.75], [25.0, 200.0]], np.float32)
    pts = np.concatenate([pts, extra, rng.uniform(5, 235, (20, 2)).astype(np.float32)])
    nxt, st, err = cv2.calcOpticalFlowPyrLK(cur, prev, pts, None, winSize=(21, 21), maxLevel=3,
                                            criteria=(cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01))
    lv = cur
    sums = []
    for _ in range(3):
        lv = cv2.pyrDown(lv)
        sums.append(int(lv.astype(np.uint64).sum()))
    np.savez_compressed(os.path.join(HERE, 'lk_320x240.npz'), cur=cur, prev=prev, pts=pts, tracked=nxt.reshape(-1, 2), status=st.reshape(-1),
                        pyr3=lv, pyr_sums=np.array(sums, np.uint64), cv2_version=np.array(cv2.__version__))
    print('points', len(pts), 'tracked ok', int(st.sum()), 'mean |flow|', np.abs(nxt.reshape(-1, 2) - pts).mean(0))


if __name__ == '__main__':
    main()
