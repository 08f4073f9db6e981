"""The Eigen stand-in that oracle/_ref/liboptimizer_ref.so (the reference's Optimizer.cc + g2o compiled unmodified) is built on, checked against numpy: the pin of
Optimizer::PoseOptimization is only as good as this header (oracle/g2o_shim/Eigen).  Linear solves (LLT, pivoted LDLT, LU), inverse / determinant (cofactor path
up to 3x3, elimination beyond), products with blocks / transposes / Map, quaternion <-> rotation-matrix conversions on every branch, q * v, quaternion products,
the fixed-size J^T W J accumulation of g2o's edges, symmetric eigenvalues.  No device needed."""
import ctypes as C
import os

import numpy as np
import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'libminieigen_selftest.so')
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason='oracle/libminieigen_selftest.so not built')
v = C.c_void_p


def _p(a):
    return a.ctypes.data_as(v)


def _lib():
    L = C.CDLL(LIB)
    L.me_solve.restype = C.c_int
    L.me_inverse_det.restype = C.c_double
    return L


def _rot(rng, angle=None):
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    th = rng.uniform(-np.pi, np.pi) if angle is None else angle
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def test_linear_solves_inverse_and_determinant():
    L = _lib(); rng = np.random.RandomState(0)
    for n in (1, 2, 3, 4, 6, 9, 17):
        for trial in range(5):
            M = rng.normal(size=(n, n)); A = M @ M.T + n * np.eye(n) * (10.0 ** -trial)       # SPD, increasingly ill-conditioned
            b = rng.normal(size=n)
            xs = [np.zeros(n) for _ in range(3)]
            pos = L.me_solve(n, _p(np.ascontiguousarray(A)), _p(b), *[_p(x) for x in xs])
            ref = np.linalg.solve(A, b)
            tol = 1e-10 * max(1.0, np.linalg.cond(A))
            assert pos == 1
            for x in xs:
                assert np.abs(x - ref).max() <= tol * max(1.0, np.abs(ref).max())
            G = rng.normal(size=(n, n)) + 2 * np.eye(n)                                            # general matrix: inverse / determinant
            inv = np.zeros((n, n)); det = L.me_inverse_det(n, _p(np.ascontiguousarray(G)), _p(inv))
            assert np.allclose(inv, np.linalg.inv(G), rtol=1e-9, atol=1e-9 * np.abs(np.linalg.inv(G)).max())
            assert np.isclose(det, np.linalg.det(G), rtol=1e-9, atol=1e-12)
    S = np.diag([3.0, -2.0, 1.0])                                                                  # indefinite: LDLT reports it, still solves
    xs = [np.zeros(3) for _ in range(3)]
    assert L.me_solve(3, _p(S), _p(np.ones(3)), *[_p(x) for x in xs]) == 0
    assert np.allclose(xs[1], [1 / 3, -0.5, 1.0]) and np.allclose(xs[2], [1 / 3, -0.5, 1.0])
    A3 = rng.normal(size=(3, 3)); inv3 = np.zeros((3, 3)); d3 = C.c_double()
    L.me_inverse3(_p(np.ascontiguousarray(A3)), _p(inv3), C.byref(d3))
    assert np.allclose(inv3, np.linalg.inv(A3), rtol=1e-10, atol=1e-12) and np.isclose(d3.value, np.linalg.det(A3), rtol=1e-12)


def test_products_blocks_transposes_maps():
    L = _lib(); rng = np.random.RandomState(1)
    for r, k, c in ((1, 1, 1), (2, 6, 6), (6, 2, 1), (5, 7, 3), (12, 1, 9)):
        A, B, D = rng.normal(size=(r, k)), rng.normal(size=(k, c)), rng.normal(size=(c, r))
        out = np.zeros((r, c))
        L.me_gemm(r, k, c, _p(A), _p(B), _p(D), _p(out))
        assert np.allclose(out, A @ B + D.T, rtol=1e-13, atol=1e-13)
    J, W, e = rng.normal(size=(2, 6)), np.diag(rng.uniform(0.5, 2, 2)), rng.normal(size=2)
    H = np.zeros((6, 6)); b = np.zeros(6)
    L.me_quadratic_form(_p(J), _p(np.ascontiguousarray(W)), _p(e), C.c_double(0.25), _p(H), _p(b))
    assert np.allclose(H, J.T @ W @ J + 0.25 * np.eye(6), rtol=1e-13, atol=1e-13) and np.allclose(b, -J.T @ W @ e, rtol=1e-13, atol=1e-13)


def test_quaternions_on_every_branch():
    L = _lib(); rng = np.random.RandomState(2)
    # trace > 0 and the three largest-diagonal branches (rotations by ~pi about x, y, z), plus random ones
    specials = [np.eye(3), np.diag([1.0, -1.0, -1.0]), np.diag([-1.0, 1.0, -1.0]), np.diag([-1.0, -1.0, 1.0])]
    near_pi = []
    for ax in np.eye(3):
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]]); th = np.pi - 1e-3
        near_pi.append(np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K)
    for R in specials + near_pi + [_rot(rng) for _ in range(200)]:
        R2 = _rot(rng); vec = rng.normal(size=3)
        q = np.zeros(4); Rb = np.zeros((3, 3)); qv = np.zeros(3); R12 = np.zeros((3, 3))
        L.me_quaternion(_p(np.ascontiguousarray(R)), _p(vec), _p(np.ascontiguousarray(R2)), _p(q), _p(Rb), _p(qv), _p(R12))
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert np.allclose(Rb, R, atol=1e-12) and np.allclose(qv, R @ vec, atol=1e-12) and np.allclose(R12, R @ R2, atol=1e-12)


def test_symmetric_eigenvalues():
    L = _lib(); rng = np.random.RandomState(3)
    for n in (1, 2, 3, 6, 10):
        M = rng.normal(size=(n, n)); A = M + M.T
        ev = np.zeros(n); L.me_eigenvalues(n, _p(np.ascontiguousarray(A)), _p(ev))
        assert np.allclose(ev, np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-10)
