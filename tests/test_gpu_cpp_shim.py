"""The reference-shaped C++ headers (include/sgslam/*.h: ORBextractor, ORBmatcher, RmDynamicPointsGeometry) compiled without
OpenCV and run on the GPU; the expected results are the CPU oracle's."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import scenarios as S
from pysgs import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    exe = os.path.join(ROOT, 'tests', 'cpp', 'test_shim')
    src = os.path.join(ROOT, 'tests', 'cpp', 'test_shim.cpp')
    lib = os.path.join(ROOT, 'sg-slam_b200', 'lib')
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-I', os.path.join(ROOT, 'include'), src, '-o', exe, '-L', lib, '-lsgs_cuda', '-Wl,-rpath,' + lib, '-ldl', '-lpthread', '-lrt'])
    return exe


def test_shim_compiles_without_gpu():
    _build()


@pytest.mark.gpu
def test_shim_on_gpu(tmp_path):
    exe = _build()
    img = synth.frame_s1(640, 480, 21)
    kps, desc = O.extract(img)
    s = S.random_lastframe_scenario(3, n_cur=900, n_last=1000)
    fo = O.FrameArrays(s['kps'], s['uright'], s['desc'], 640, 480, s['cam']['fx'], s['cam']['fy'], s['cam']['cx'], s['cam']['cy'], s['cam']['bf'], s['sf'])
    nm, mp, _ = O.search_by_projection_last(fo, s['Tcw_cur'], s['Tcw_last'], s['last_has'], s['last_xyz'], s['last_desc'], s['last_obs'], s['last_oct'], s['last_angle'], 15.0)
    d = S.dynreject_scenario(5, n=len(kps))
    dk = kps.copy(); dk['x'] = d['cur'][:, 0]; dk['y'] = d['cur'][:, 1]
    nkeep, keep, _, restored = O.dynreject(d['cur'], d['prev'], d['F'], d['boxes'], True, 1000)
    assert not restored
    path = tmp_path / 'scenario.bin'
    with open(path, 'wb') as f:
        np.array([640, 480, len(kps), len(s['kps']), len(s['last_has']), nm, nkeep, len(dk)], np.int32).tofile(f)
        img.tofile(f); kps.tofile(f); desc.tofile(f)
        s['kps'].tofile(f); s['uright'].astype(np.float32).tofile(f); s['desc'].tofile(f); s['sf'].astype(np.float32).tofile(f)
        s['Tcw_cur'].astype(np.float32).tofile(f); s['Tcw_last'].astype(np.float32).tofile(f)
        s['last_has'].astype(np.uint8).tofile(f); s['last_obs'].astype(np.uint8).tofile(f); s['last_xyz'].astype(np.float32).tofile(f)
        s['last_desc'].tofile(f); s['last_oct'].astype(np.int32).tofile(f); s['last_angle'].astype(np.float32).tofile(f); mp.astype(np.int32).tofile(f)
        dk.tofile(f); desc.tofile(f); d['prev'].astype(np.float32).tofile(f); d['F'].astype(np.float64).tofile(f)
        d['boxes'].astype(np.float32).tofile(f); keep.astype(np.uint8).tofile(f)
        # 4. LK + findFundamentalMat shims: a second frame of the stream, the oracle's tracks and its F on them
        frames, _ = synth.stream_s2(2, 640, 480, seed=9)
        k2, _ = O.extract(frames[1])
        pts = np.stack([k2['x'], k2['y']], 1).astype(np.float32)
        trk = O.lk_track(frames[1], frames[0], pts)
        Fo, _, _ = O.find_fundamental_ransac(pts, trk)
        assert Fo is not None
        np.array([len(pts)], np.int32).tofile(f); frames[1].tofile(f); frames[0].tofile(f); pts.tofile(f); trk.astype(np.float32).tofile(f)
        Fo.astype(np.float64).tofile(f)
    out = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'OK shim' in out.stdout
