"""Pin the CPU oracle (oracle/sgs_oracle.cpp) against the golden vectors generated with the REAL OpenCV
primitives by tests/golden/make_golden.py, and -- when cv2 is importable -- against cv2 live.

Bar: bit-exact (integer / byte / index work; float fields compared by their bit patterns)."""
import os

import numpy as np
import pytest

import oracle as O

CASES = ['s1_640x480', 's1_320x240', 'noise_200x160']


def _xor_checksum(p):
    return int(np.bitwise_xor.reduce(p.reshape(-1).astype(np.uint64) * (np.arange(p.size, dtype=np.uint64) % 65521 + 1)))


@pytest.mark.parametrize('name', CASES)
def test_extract_matches_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, 'extract_%s.npz' % name))
    nfeat, nlev, ini, mn = [int(v) for v in g['params']]
    p = O.params(nfeat, 1.2, nlev, ini, mn)
    t = O.orb_tables(p)
    assert t['nPerLevel'].tolist() == g['per'].tolist()
    assert t['umax'].tolist() == g['umax'].tolist()
    assert t['scale'].view(np.uint32).tolist() == g['scale'].view(np.uint32).tolist()
    d = O.ExtractDump(g['image'], p)
    # pyramid: sizes + checksums for every level, full image for the last level
    for lvl in range(nlev):
        assert list(d.pyramid[lvl].shape[::-1]) == g['level_sizes'][lvl].tolist()
        assert int(d.pyramid[lvl].astype(np.uint64).sum()) == int(g['pyr_sums'][lvl])
        assert _xor_checksum(d.pyramid[lvl]) == int(g['pyr_xor'][lvl])
        assert len(d.cands[lvl]) == int(g['ncands'][lvl])
        if d.blurred[lvl] is not None:
            assert int(d.blurred[lvl].astype(np.uint64).sum()) == int(g['blur_sums'][lvl])
    assert np.array_equal(d.pyramid[nlev - 1], g['pyr_last'])
    assert np.array_equal(d.blurred[nlev - 1], g['blur_last'])
    # candidates in reference order (level 0 and last level)
    assert np.array_equal(d.cands[0], g['cands_l0'])
    assert np.array_equal(d.cands[nlev - 1], g['cands_last'])
    # final keypoints (bitwise) and descriptors
    assert len(d.kps) == len(g['kps'])
    assert d.kps.tobytes() == g['kps'].tobytes()
    assert np.array_equal(d.desc, g['desc'])
    # the plain entry point agrees with the dump entry point
    k2, d2 = O.extract(g['image'], p)
    assert k2.tobytes() == d.kps.tobytes() and np.array_equal(d2, d.desc)


def test_primitives_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'primitives.npz'))
    at = np.array([O.fast_atan2(y, x) for y, x in zip(g['atan_y'], g['atan_x'])], np.float32)
    assert at.view(np.uint32).tolist() == g['atan'].view(np.uint32).tolist()
    assert np.array_equal(O.resize(g['small'], 44, 31), g['resize_44x31'])
    assert np.array_equal(O.blur(g['small']), g['blur'])
    k = O.fast_view(g['small'], 20, True).astype(np.float32)
    assert np.array_equal(k, g['fast20'])


def test_pattern_taps_stay_inside_patch():
    # rotated BRIEF taps reach at most 18 px from the keypoint (pattern corner (-13,-13) -> radius 18.4); keypoints are
    # >= 19 px (EDGE_THRESHOLD) from the level border (ORBextractor.cc:774-777), so taps never leave the level image and the
    # pyramid's 19-px border is never read -- but the blur's reflect-101 border IS felt by taps within 3 px of the edge.
    ext = max(O.pattern_extent(a) for a in np.arange(0, 360, 0.25))
    assert 15 < ext <= 18


cv2 = pytest.importorskip('cv2') if os.environ.get('SGS_SKIP_CV2') is None else None


@pytest.mark.skipif(cv2 is None, reason='cv2 not importable')
def test_primitives_against_cv2_live():
    rng = np.random.RandomState(123)
    # resize: ORB-SLAM level chain sizes for 640x480 and 1280x720 plus odd sizes
    for (sw, sh, dw, dh) in [(640, 480, 533, 400), (533, 400, 444, 333), (214, 161, 179, 134), (1280, 720, 1067, 600),
                             (97, 61, 81, 51), (50, 40, 49, 39), (33, 200, 28, 167)]:
        src = rng.randint(0, 256, (sh, sw)).astype(np.uint8)
        assert np.array_equal(O.resize(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)), (sw, sh, dw, dh)
    # blur
    for (w, h) in [(64, 80), (179, 134), (7, 9), (31, 8)]:
        src = rng.randint(0, 256, (h, w)).astype(np.uint8)
        assert np.array_equal(O.blur(src), cv2.GaussianBlur(src, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)), (w, h)
    # FAST on views, both thresholds, NMS on/off; smooth-ish image so that corners are sparse, plus pure noise
    base = rng.randint(0, 256, (120, 160)).astype(np.uint8)
    smooth = cv2.GaussianBlur(base, (5, 5), 1.2)
    for img in (base, smooth):
        for thr in (7, 20):
            for nms in (True, False):
                for (x0, y0, x1, y1) in [(0, 0, 160, 120), (16, 16, 53, 54), (40, 30, 77, 67), (100, 80, 160, 120)]:
                    view = np.ascontiguousarray(img[y0:y1, x0:x1])
                    det = cv2.FastFeatureDetector_create(thr, nms, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                    ref = np.array([[k.pt[0], k.pt[1], k.response] for k in det.detect(view)], np.float32).reshape(-1, 3)
                    got = O.fast_view(view, thr, nms).astype(np.float32)
                    if not nms:
                        got[:, 2] = 0  # cv2 reports response 0 without NMS
                    assert np.array_equal(got, ref), (thr, nms, x0, y0)
    # fastAtan2
    ys = rng.randint(-100000, 100000, 5000); xs = rng.randint(-100000, 100000, 5000)
    for y, x in zip(ys, xs):
        assert np.float32(O.fast_atan2(y, x)).view(np.uint32) == np.float32(cv2.fastAtan2(float(y), float(x))).view(np.uint32)
