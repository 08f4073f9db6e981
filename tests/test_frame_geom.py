"""Oracle of Frame::isInFrustum / ComputeStereoFromRGBD against golden vectors made with the real cv2 matrix primitives
(tests/golden/make_golden_frustum.py).  Floats must be bit-identical; the predicted level may differ only where
log(ratio)/log(scaleFactor) is within 1e-5 of an integer (logf implementations differ in the last ulp)."""
import os

import numpy as np

import oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'frustum.npz'))


def check_frustum(out, g=G):
    assert np.array_equal(out['inview'], g['inview'])
    for k in ('proj_x', 'proj_y', 'proj_xr', 'view_cos'):
        assert out[k].tobytes() == g[k].tobytes(), k
    diff = np.nonzero(out['level'] != g['level'])[0]
    arg = g['level_arg'][diff]
    assert len(diff) <= 2 and np.all(np.abs(arg - np.round(arg)) < 1e-5), (diff, arg)


def test_is_in_frustum_matches_cv2_primitives():
    out = O.is_in_frustum(G['Tcw'], G['cam'], 8, float(G['logsf']), G['xyz'], G['normal'], G['min_dist'], G['max_dist'], 0.5)
    assert 300 < out['inview'].sum() < 3500
    check_frustum(out)


def test_stereo_from_rgbd():
    k = np.zeros(len(G['kx']), O.KP_DTYPE); k['x'] = G['kx']; k['y'] = G['ky']
    ur, dz = O.stereo_from_rgbd(k, G['depth'], 40.0)
    assert ur.tobytes() == G['u_right'].tobytes() and dz.tobytes() == G['depth_out'].tobytes()
    assert (ur == -1).sum() > 10


def test_undistort_points_bit_exact_vs_cv2():
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'undistort.npz'))
    for name in ('TUM1', 'TUM2'):
        K = g[name + '_K']
        got = O.undistort_points(g[name + '_pts'], K[0], K[1], K[2], K[3], g[name + '_D'])
        assert got.tobytes() == g[name + '_und'].tobytes(), name
