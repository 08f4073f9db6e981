"""Oracle of Optimizer::PoseOptimization (src/Optimizer.cc:239-451).  The restatement is pinned against the reference's own Optimizer.cc + g2o in
tests/test_optimizer_ref.py; checked here independently of g2o: the result is the least-squares optimum of the final inlier set (an independent numpy Gauss-Newton on rotation
matrices reaches the same pose), gross outliers are flagged, clean data keeps every edge, degenerate inputs behave like the reference."""
import numpy as np

import oracle as O
import scenarios as S


def _run(s):
    c = s['cam']
    return O.pose_optimization(s['T0'], s['has'], s['xyz'], s['xy'], s['octave'], s['uright'], s['inv_s2'], c['fx'], c['fy'], c['cx'], c['cy'], c['bf'])


def _gauss_newton(s, T_start, sel, iters=15):
    """Plain Gauss-Newton on [R|t] with a left-multiplied rotation-vector update (different parametrisation code path from the oracle's)."""
    c = s['cam']
    R = T_start[:3, :3].astype(np.float64); t = T_start[:3, 3].astype(np.float64)
    X = s['xyz'][sel].astype(np.float64); xy = s['xy'][sel].astype(np.float64); ur = s['uright'][sel].astype(np.float64)
    w = s['inv_s2'][s['octave'][sel]].astype(np.float64)
    for _ in range(iters):
        P = X @ R.T + t
        x, y, z = P[:, 0], P[:, 1], P[:, 2]
        iz = 1 / z
        stereo = ur >= 0
        r = np.stack([xy[:, 0] - (c['fx'] * x * iz + c['cx']), xy[:, 1] - (c['fy'] * y * iz + c['cy']), np.where(stereo, ur - (c['fx'] * x * iz + c['cx'] - c['bf'] * iz), 0.0)], 1)
        J = np.zeros((len(X), 3, 6))
        J[:, 0] = np.stack([x * y * iz * iz * c['fx'], -(1 + x * x * iz * iz) * c['fx'], y * iz * c['fx'], -iz * c['fx'], 0 * x, x * iz * iz * c['fx']], 1)
        J[:, 1] = np.stack([(1 + y * y * iz * iz) * c['fy'], -x * y * iz * iz * c['fy'], -x * iz * c['fy'], 0 * x, -iz * c['fy'], y * iz * iz * c['fy']], 1)
        J[:, 2] = J[:, 0] + np.stack([-c['bf'] * y * iz * iz, c['bf'] * x * iz * iz, 0 * x, 0 * x, 0 * x, -c['bf'] * iz * iz], 1)
        J[~stereo, 2] = 0
        H = np.einsum('nij,n,nik->jk', J, w, J); b = -np.einsum('nij,n,ni->j', J, w, r)
        dx = np.linalg.solve(H, b)
        om, up = dx[:3], dx[3:]
        th = np.linalg.norm(om)
        K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        if th < 1e-12:
            dR, V = np.eye(3) + K, np.eye(3)
        else:
            dR = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
            V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K
        R = dR @ R; t = dR @ t + V @ up
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return T


def test_recovers_pose_and_flags_gross_outliers():
    for seed in (1, 2, 3):
        s = S.pose_scenario(seed)
        nin, T, outl = _run(s)
        used = s['has'] == 1
        assert np.abs(T - s['T_true']).max() < 2e-3 and np.abs(s['T0'] - s['T_true']).max() > 2e-2
        assert outl[used & s['gross']].mean() > 0.9 and outl[used & ~s['gross']].mean() < 0.1
        assert nin == int(used.sum()) - int(outl[used].sum())
        # the pose is the least-squares optimum of the edges that entered the last round (= the inliers of round 3, which the final flags
        # reproduce except for borderline edges): an independent Gauss-Newton from the same start lands on it
        Tg = _gauss_newton(s, s['T0'], used & (outl == 0))
        assert np.abs(Tg - T).max() < 5e-5, np.abs(Tg - T).max()


def test_clean_data_and_degenerate_inputs():
    s = S.pose_scenario(5, outlier_frac=0.0, noise=0.2)
    nin, T, outl = _run(s)
    assert nin >= int(s['has'].sum()) - 3 and np.abs(T - s['T_true']).max() < 1e-3
    s2 = dict(s); s2['has'] = np.zeros_like(s['has']); s2['has'][:2] = 1
    nin, T, outl = _run(s2)
    assert nin == 0 and np.array_equal(T, s['T0'])                     # fewer than 3 correspondences: untouched (src/Optimizer.cc:338-339)
    s3 = dict(s); s3['has'] = np.zeros_like(s['has']); s3['has'][:8] = 1
    nin, T, outl = _run(s3)                                            # fewer than 10 edges: a single round (:427-428)
    assert 0 <= nin <= 8
