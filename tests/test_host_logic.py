"""Host-side logic of the product (planning + the quadtree core shared between host and device) against the CPU oracle.
Runs without a GPU: sg-slam_b200/csrc/host_checks.cpp compiles the same headers the CUDA kernels use with g++."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'sg-slam_b200', 'csrc')


@pytest.fixture(scope='module')
def hc():
    subprocess.check_call(['make', '-C', CSRC, '-s', 'hostcheck'])
    return C.CDLL(os.path.join(ROOT, 'sg-slam_b200', 'lib', 'libsgs_hostcheck.so'))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize('w,h,nfeat,nlev', [(640, 480, 1000, 8), (1280, 720, 2000, 8), (320, 240, 500, 8), (200, 160, 300, 4), (752, 480, 1200, 8)])
def test_plan_matches_oracle(hc, w, h, nfeat, nlev):
    p = O.params(nfeat, 1.2, nlev, 20, 7)
    wh = np.zeros((nlev, 2), np.int32); npl = np.zeros(nlev, np.int32); umax = np.zeros(16, np.int32)
    sc = np.zeros(nlev, np.float32); nc = C.c_int32(); geom = np.zeros((nlev, 8), np.int32)
    assert hc.sgs_hostcheck_plan(C.byref(p), w, h, _p(wh), _p(npl), _p(umax), _p(sc), C.byref(nc), _p(geom)) == 0
    t = O.orb_tables(p)
    assert npl.tolist() == t['nPerLevel'].tolist()
    assert umax.tolist() == t['umax'].tolist()
    assert sc.view(np.uint32).tolist() == t['scale'].view(np.uint32).tolist()
    for l in range(nlev):
        assert tuple(wh[l]) == O.level_size(p, w, h, l)


def _run_qt(hc, p, w, h, level, cands_xyz, n_override=-1):
    c = np.ascontiguousarray(cands_xyz, np.int32)
    out = np.zeros((len(c) + 64, 3), np.int32)
    n = hc.sgs_hostcheck_quadtree(C.byref(p), w, h, level, _p(c), len(c), n_override, _p(out), len(out))
    assert n >= 0, n
    return out[:n]


def _oracle_qt(p, w, h, level, cands_xyz, N):
    lw, lh = O.level_size(p, w, h, level)
    sel = O.octree(np.asarray(cands_xyz, np.float32), 16, lw - 16, 16, lh - 16, N)
    return np.asarray(cands_xyz, np.int32)[sel]


def test_quadtree_core_on_golden_candidates(hc, golden_dir):
    for name in ['s1_640x480', 's1_320x240', 'noise_200x160']:
        g = np.load(os.path.join(golden_dir, 'extract_%s.npz' % name))
        nfeat, nlev, ini, mn = [int(v) for v in g['params']]
        p = O.params(nfeat, 1.2, nlev, ini, mn)
        img = g['image']
        d = O.ExtractDump(img, p)
        t = O.orb_tables(p)
        for lvl in range(nlev):
            c = d.cands[lvl]
            got = _run_qt(hc, p, img.shape[1], img.shape[0], lvl, c.astype(np.int32))
            ref = _oracle_qt(p, img.shape[1], img.shape[0], lvl, c, int(t['nPerLevel'][lvl]))
            assert np.array_equal(got, ref), (name, lvl)
            # candidate order must not matter to the product (it sorts); shuffle and repeat
            perm = np.random.RandomState(lvl).permutation(len(c))
            got2 = _run_qt(hc, p, img.shape[1], img.shape[0], lvl, c[perm].astype(np.int32))
            assert np.array_equal(got2, ref), (name, lvl, 'shuffled')


@pytest.mark.parametrize('seed', range(12))
def test_quadtree_core_random(hc, seed):
    """Random candidate sets (distinct pixels, many equal responses to exercise the tie-break), various targets N,
    including N=0, N=1, N > number of candidates, and a 2-root (wide) geometry."""
    rng = np.random.RandomState(seed)
    w, h = [(640, 480), (1280, 720), (400, 200), (330, 300)][seed % 4]
    p = O.params(1000, 1.2, 1, 20, 7)
    iw, ih = w - 32, h - 32
    # only pixels FAST can report: 3 <= x < iw-3
    n = [0, 1, 2, 7, 50, 400, 3000, 9000, 100, 30, 1500, 5][seed]
    xs = rng.randint(3, iw - 3, 4 * n + 8); ys = rng.randint(3, ih - 3, 4 * n + 8)
    pix = np.unique(np.stack([ys, xs], 1), axis=0)
    pix = pix[rng.permutation(len(pix))[:n]]
    # reference order: cells row-major then pixels row-major -> build via oracle-like key so that 'first wins' is meaningful
    sc = rng.randint(7, 12, len(pix))  # few distinct values => many ties
    c = np.stack([pix[:, 1], pix[:, 0], sc], 1).astype(np.int32).reshape(-1, 3)
    # sort into reference candidate order
    hcgeom = np.zeros((1, 8), np.int32); wh = np.zeros((1, 2), np.int32); npl = np.zeros(1, np.int32); um = np.zeros(16, np.int32)
    scl = np.zeros(1, np.float32); nc = C.c_int32()
    assert hc.sgs_hostcheck_plan(C.byref(p), w, h, _p(wh), _p(npl), _p(um), _p(scl), C.byref(nc), _p(hcgeom)) == 0
    n_cols, n_rows, w_cell, h_cell = [int(v) for v in hcgeom[0, :4]]
    if len(c):
        ci = (c[:, 1] - 3) // h_cell; cj = (c[:, 0] - 3) // w_cell
        order = np.lexsort((c[:, 0], c[:, 1], cj, ci))
        c = c[order]
    for N in [0, 1, 5, 97, 217, 1000, 20000]:
        got = _run_qt(hc, p, w, h, 0, c, N)
        ref = _oracle_qt(p, w, h, 0, c, N)
        assert np.array_equal(got, ref), (seed, N, len(c))


def test_logf_restatement_equals_libm(hc):
    """sgs_logf.h (the device's logf for MapPoint::PredictScale, src/MapPoint.cc:402-418) against the running libm, bit for bit: a stride-61
    sweep over every positive finite float (35 M values), every float in [0.25, 16] (the range of mfMaxDistance / dist), the special cases."""
    hc.sgs_hostcheck_logf_mismatches.restype = C.c_longlong
    hc.sgs_hostcheck_logf.restype = C.c_float
    assert hc.sgs_hostcheck_logf_mismatches(C.c_uint32(1), C.c_uint32(61), C.c_longlong(0x7f800000 // 61)) == 0
    lo = int(np.float32(0.25).view(np.uint32)); hi = int(np.float32(16.0).view(np.uint32))
    assert hc.sgs_hostcheck_logf_mismatches(C.c_uint32(lo), C.c_uint32(1), C.c_longlong(hi - lo + 1)) == 0
    assert hc.sgs_hostcheck_logf(C.c_float(1.0)) == 0.0
    assert hc.sgs_hostcheck_logf(C.c_float(0.0)) == -np.inf and hc.sgs_hostcheck_logf(C.c_float(np.inf)) == np.inf
    assert np.isnan(hc.sgs_hostcheck_logf(C.c_float(-1.0))) and np.isnan(hc.sgs_hostcheck_logf(C.c_float(np.nan)))
    assert hc.sgs_hostcheck_logf_mismatches(C.c_uint32(1), C.c_uint32(1), C.c_longlong(0x00800000)) == 0       # all subnormals
