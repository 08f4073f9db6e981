#!/usr/bin/env python3
"""Golden vectors for findFundamentalMat (src/Frame.cc:469-472): the REAL cv2.findFundamentalMat(FM_RANSAC, 1.0, 0.99) and
cv2.findFundamentalMat(FM_7POINT) (= run7Point) on seeded synthetic two-view correspondences.
Run in the build container (needs cv2):  python tests/golden/make_golden_fm.py"""
import math
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def two_view(rs, n, noise, outlier_frac, outlier_mag=30.0, ang=0.02, t=(0.05, 0.01, 0.02)):
    X = np.c_[rs.uniform(-2, 2, n), rs.uniform(-1.5, 1.5, n), rs.uniform(2, 6, n)]
    K = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1]])
    R = np.array([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
    p1 = (K @ X.T).T; p1 = p1[:, :2] / p1[:, 2:]
    X2 = (R @ X.T).T + np.array(t); p2 = (K @ X2.T).T; p2 = p2[:, :2] / p2[:, 2:]
    p1 = p1 + rs.normal(0, noise, p1.shape); p2 = p2 + rs.normal(0, noise, p2.shape)
    nout = int(n * outlier_frac)
    p2[:nout] += rs.uniform(-outlier_mag, outlier_mag, (nout, 2))
    perm = rs.permutation(n)
    return p1[perm].astype(np.float32), p2[perm].astype(np.float32)


def main():
    rs = np.random.RandomState(20240924)
    out = {}
    cases = [(15, 0.2, 0.0), (40, 0.3, 0.2), (150, 0.3, 0.2), (400, 0.1, 0.1), (1000, 0.3, 0.3), (1000, 0.05, 0.02), (2000, 0.5, 0.45), (700, 0.2, 0.6)]
    for i, (n, noise, of) in enumerate(cases):
        m1, m2 = two_view(rs, n, noise, of)
        F, mask = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, 1.0, 0.99)
        out[f'r{i}_m1'] = m1; out[f'r{i}_m2'] = m2
        out[f'r{i}_F'] = np.zeros((0, 3)) if F is None else F
        out[f'r{i}_mask'] = np.zeros(0, np.uint8) if mask is None else mask.ravel().astype(np.uint8)
        print('ransac', n, noise, of, None if F is None else (F.shape, int(mask.sum())))
    for i in range(12):
        m1, m2 = two_view(rs, 7, 0.5, 0.0)
        F, _ = cv2.findFundamentalMat(m1, m2, cv2.FM_7POINT)
        out[f's{i}_m1'] = m1; out[f's{i}_m2'] = m2; out[f's{i}_F'] = np.zeros((0, 3)) if F is None else F
    out['n_ransac'] = np.array(len(cases)); out['n_seven'] = np.array(12); out['cv2_version'] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, 'fm_ransac.npz'), **out)


if __name__ == '__main__':
    main()
