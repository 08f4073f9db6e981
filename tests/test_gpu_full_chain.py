"""End-to-end identity of the GPU chain with the PURE oracle chain ("never measured"): extract -> LK -> findFundamentalMat -> dyn-reject ->
SearchByProjection(cur, last) on S2 streams, the GPU through the C ABI (sgs_tracker_extract + sgs_tracker_track_lk), the oracle through oracle/chain.cpp with
its OWN LK and its OWN F.  Extraction must be bit-exact; LK agrees to ~1e-4 px (the GPU sums the 21x21 window exactly, OpenCV in float order), which can move
a keypoint's epipolar distance across its 0.2 / 1.0 px threshold: the keep-set symmetric difference and the match-index difference per frame are measured and
bounded.  With the S2 person box active (thresholds 0.2 px inside the box, ~210 of ~1006 keypoints rejected per frame) a sub-0.03 px LK difference can also
hand RANSAC a different winning sample on a few frames, which then flips several verdicts at once.  Stated bound: mean <= 4 keypoints per frame (0.4 %),
max <= 60 on any frame (6 %), at least half of the frames identical end to end; observed on B200: mean 1.5 / max 27 with the box, 0.05 / 3 without
(bench.py, 256 frames, detector boxes)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as O  # noqa: E402
from pysgs import binding as B  # noqa: E402
from pysgs import synth  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

W, H, NF, TH = 640, 480, 1000, 15.0


def test_gpu_chain_against_the_pure_oracle_chain():
    import bench
    nb, unique = 64, 32
    frames, boxes, unique = bench.make_frames(nb, 7, W, H, unique=unique)
    pidx = bench.prev_index(nb, unique)
    camd = dict(synth.TUM3)
    sf = synth.scale_factors(); cam = B.make_camera(W, H, camd, sf)
    trk = B.Tracker(W, H, cam, NF, 1.2, 8, 20, 7, max_batch=nb, point_cap=NF + 64, max_boxes=4, device=0)
    cap, pcap = trk.cap, trk.point_cap
    L, v = B.lib(), C.c_void_p
    P = lambda a: a.ctypes.data_as(v)
    kps = np.zeros((nb, cap), B.KP_DTYPE); desc = np.zeros((nb, cap, 32), np.uint8); n = np.zeros(nb, np.int32)
    B.check(L.sgs_tracker_extract(trk.h, P(frames), nb, C.c_size_t(W * H), W, P(kps), P(desc), cap, P(n)))
    ti = bench.make_track_inputs(kps, desc, n, boxes, cap, pcap, pidx, W, H, camd)
    o = dict(kps=np.zeros((nb, cap), B.KP_DTYPE), desc=np.zeros((nb, cap, 32), np.uint8), ur=np.zeros((nb, cap), np.float32), cnt=np.zeros(nb, np.int32),
             mp=np.zeros((nb, cap), np.int32), nm=np.zeros(nb, np.int32))
    B.check(L.sgs_tracker_track_lk(trk.h, nb, P(ti['pidx']), P(ti['ur']), v(0), P(ti['boxes']), P(ti['nb']), P(ti['have']), P(ti['lxyz']), P(ti['ldesc']), P(ti['lflags']),
                                   P(ti['loct']), P(ti['lang']), P(ti['ln']), P(ti['T']), P(ti['T']), C.c_float(TH), 0, 1, P(o['kps']), P(o['desc']), P(o['ur']), P(o['cnt']),
                                   P(o['mp']), P(o['nm'])))
    trk.close()
    ch = O.Chain(frames, pidx, ti, camd, cap, nfeatures=NF, th=TH, want_outputs=True)
    ch.run(0, nb, nthreads=O.online_cpus())
    r = ch.out
    keep_diff, match_diff, in_box_rejected = [], [], 0
    for f in range(nb):
        m = int(r['counts'][f])
        assert m == n[f] and r['kps'][f, :m].tobytes() == kps[f, :m].tobytes() and np.array_equal(r['desc'][f, :m], desc[f, :m]), 'extraction must be bit-exact'
        ko = np.ones(m, bool) if r['restored'][f] else r['keep'][f, :m].astype(bool)
        ng = int(o['cnt'][f])
        kg = np.zeros(m, bool)
        j = 0
        for i in range(m):                                   # the survivors are an ordered subsequence of the extracted keypoints
            if j < ng and kps[f, i] == o['kps'][f, j]:
                kg[i] = True; j += 1
        assert j == ng
        keep_diff.append(int((kg != ko).sum()))
        mo = np.full(m, -1, np.int64); mo[np.nonzero(ko)[0]] = r['match'][f, :int(ko.sum())]
        mg = np.full(m, -1, np.int64); mg[np.nonzero(kg)[0]] = o['mp'][f, :ng]
        match_diff.append(int((mo != mg).sum()))
        in_box_rejected += int((~ko).sum())
    keep_diff, match_diff = np.array(keep_diff), np.array(match_diff)
    print('keep-set symmetric difference per frame: mean %.3f max %d; match-index difference: mean %.3f max %d; identical frames %d / %d; oracle rejects %.1f keypoints per frame'
          % (keep_diff.mean(), keep_diff.max(), match_diff.mean(), match_diff.max(), int(((keep_diff == 0) & (match_diff == 0)).sum()), nb, in_box_rejected / nb))
    assert in_box_rejected > nb, 'the scenario must actually reject keypoints (person box + epipolar test)'
    assert keep_diff.mean() <= 4.0 and keep_diff.max() <= 60
    assert match_diff.mean() <= 4.0 and match_diff.max() <= 60
    assert int(((keep_diff == 0) & (match_diff == 0)).sum()) >= nb // 2
