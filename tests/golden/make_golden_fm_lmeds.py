#!/usr/bin/env python3
"""Golden vectors for the small-sample branch of cv::findFundamentalMat(FM_RANSAC, 1.0, 0.99) (src/Frame.cc:469-472): with 8..14 pairs OpenCV
runs LMedS instead of RANSAC (fundam.cpp).  Made with the REAL cv2.  Run in the build container:  python tests/golden/make_golden_fm_lmeds.py"""
import importlib.util
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('g', os.path.join(HERE, 'make_golden_fm.py')); g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)


def main():
    rs = np.random.RandomState(20240925)
    out = {}
    cases = [(14, 0.4, 0.2), (14, 1.5, 0.4), (14, 0.2, 0.0), (14, 0.8, 0.3), (14, 3.0, 0.5), (14, 0.3, 0.1),        # median = 8th error: well defined
             (8, 0.2, 0.0), (9, 0.3, 0.1), (10, 0.3, 0.2), (11, 0.5, 0.2), (12, 0.3, 0.3), (13, 0.2, 0.1)]          # median among the 7 sample points: ~1e-27 noise
    for i, (n, noise, of) in enumerate(cases):
        m1, m2 = g.two_view(rs, n, noise, of)
        F, mask = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, 1.0, 0.99)
        F2, _ = cv2.findFundamentalMat(m1, m2, cv2.FM_LMEDS, 1.0, 0.99)
        assert (F is None) == (F2 is None) and (F is None or np.array_equal(F, F2))          # the RANSAC flag really takes the LMedS path
        out[f'l{i}_m1'] = m1; out[f'l{i}_m2'] = m2
        out[f'l{i}_F'] = np.zeros((0, 3)) if F is None else F
        out[f'l{i}_mask'] = np.zeros(0, np.uint8) if mask is None else mask.ravel().astype(np.uint8)
        print(n, noise, of, None if F is None else int(mask.sum()))
    # exactly 7 pairs: cv::findFundamentalMat runs the 7-point solver itself and returns its 1..3 solutions STACKED (3k x 3); the reference reads
    # F.at<double>(0..2, 0..2) = the first (src/Frame.cc:613-627).  Fewer than 7: empty.
    for j in range(4):
        m1, m2 = g.two_view(rs, 7, 0.3, 0.0)
        F, mask = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, 1.0, 0.99)
        out[f's{j}_m1'] = m1; out[f's{j}_m2'] = m2; out[f's{j}_F'] = np.zeros((0, 3)) if F is None else F
        print(7, None if F is None else F.shape)
    m1, m2 = g.two_view(rs, 6, 0.3, 0.0)
    F, mask = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, 1.0, 0.99)
    assert F is None
    out['n_seven'] = np.array(4)
    out['n_cases'] = np.array(len(cases)); out['cv2_version'] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, 'fm_lmeds.npz'), **out)


if __name__ == '__main__':
    main()
