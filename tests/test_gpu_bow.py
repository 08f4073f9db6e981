"""GPU parity of DBoW2 transform and ORBmatcher::SearchByBoW(KeyFrame*, Frame&) through the C ABI against the CPU oracle: integer work,
bit-exact (word / node ids, weights as stored, assignments, counts)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as O  # noqa: E402
import scenarios as S  # noqa: E402
from pysgs import binding as B  # noqa: E402


class GpuVoc:
    def __init__(self, voc):
        self.h = C.c_void_p()
        v = C.c_void_p
        B.check(B.lib().sgs_vocabulary_create(0, voc['k'], voc['L'], len(voc['parent']), voc['parent'].ctypes.data_as(v), np.ascontiguousarray(voc['desc']).ctypes.data_as(v),
                                              voc['weight'].ctypes.data_as(v), C.byref(self.h)))

    def close(self):
        B.lib().sgs_vocabulary_destroy(self.h)


def _transform_gpu(gv, desc, counts, levelsup):
    import torch
    F, cap = desc.shape[:2]
    dd = torch.from_numpy(np.ascontiguousarray(desc)).cuda(); dc = torch.from_numpy(np.asarray(counts, np.int32)).cuda()
    word = torch.full((F, cap), -7, dtype=torch.int32, device='cuda'); w = torch.zeros((F, cap), dtype=torch.float64, device='cuda'); node = torch.full((F, cap), -7, dtype=torch.int32, device='cuda')
    v = C.c_void_p
    B.check(B.lib().sgs_bow_transform_batch_device(gv.h, v(dd.data_ptr()), v(dc.data_ptr()), cap, F, levelsup, v(word.data_ptr()), v(w.data_ptr()), v(node.data_ptr()), v(0)))
    torch.cuda.synchronize()
    return word, w, node


@pytest.mark.parametrize('k,L', [(10, 3), (4, 5), (33, 2)])
def test_transform(k, L):
    voc = S.random_vocabulary(5 + k, k=k, L=L)
    V = O.Vocabulary(voc['k'], voc['L'], voc['parent'], voc['desc'], voc['weight'])
    gv = GpuVoc(voc)
    try:
        sc = [S.bow_pair_scenario(s, voc, n_kf=700, n_f=1) for s in (1, 2)]
        desc = np.stack([c['kf_desc'] for c in sc]); counts = [700, 333]
        for levelsup in (0, 1, 2, 4):
            word, w, node = _transform_gpu(gv, desc, counts, levelsup)
            word, w, node = word.cpu().numpy(), w.cpu().numpy(), node.cpu().numpy()
            for f in range(2):
                n = counts[f]
                ow, owt, on = V.transform(desc[f, :n], levelsup)
                assert np.array_equal(word[f, :n], ow) and np.array_equal(node[f, :n], on) and w[f, :n].tobytes() == owt.tobytes()
                assert np.all(word[f, n:] == -7)
    finally:
        gv.close()


@pytest.mark.parametrize('seed,nk,nf,flips,nnratio', [(1, 1000, 1000, 40, 0.7), (2, 1100, 400, 80, 0.6), (3, 50, 1100, 10, 0.9), (4, 1000, 1000, 0, 0.7)])
def test_search_by_bow(seed, nk, nf, flips, nnratio):
    import torch
    voc = S.random_vocabulary(9, k=10, L=3)
    V = O.Vocabulary(voc['k'], voc['L'], voc['parent'], voc['desc'], voc['weight'])
    gv = GpuVoc(voc)
    try:
        F, cap = 3, 1100
        scs = [S.bow_pair_scenario(seed * 10 + i, voc, n_kf=nk, n_f=nf, flips=flips) for i in range(F)]
        kd = np.zeros((F, cap, 32), np.uint8); fd = np.zeros((F, cap, 32), np.uint8); ka = np.zeros((F, cap), np.float32); fa = np.zeros((F, cap), np.float32)
        kv = np.zeros((F, cap), np.uint8)
        for i, s in enumerate(scs):
            kd[i, :nk] = s['kf_desc']; fd[i, :nf] = s['f_desc']; ka[i, :nk] = s['kf_angle']; fa[i, :nf] = s['f_angle']; kv[i, :nk] = s['kf_valid']
        kn = np.array([nk, nk, max(1, nk // 2)], np.int32); fn = np.array([nf, max(1, nf // 3), nf], np.int32)
        for levelsup in (1, 2):
            _, kw, knode = _transform_gpu(gv, kd, kn, levelsup); _, fw, fnode = _transform_gpu(gv, fd, fn, levelsup)
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
            t = dict(kv=dev(kv), kd=dev(kd), ka=dev(ka), kn=dev(kn), fd=dev(fd), fa=dev(fa), fn=dev(fn))
            for ori in (0, 1):
                m = torch.zeros((F, cap), dtype=torch.int32, device='cuda'); nm = torch.zeros(F, dtype=torch.int32, device='cuda')
                a = B.BowBatch()
                a.kf_node, a.kf_weight, a.kf_valid, a.kf_desc, a.kf_angle, a.kf_n, a.kf_cap = knode.data_ptr(), kw.data_ptr(), t['kv'].data_ptr(), t['kd'].data_ptr(), t['ka'].data_ptr(), t['kn'].data_ptr(), cap
                a.f_node, a.f_weight, a.f_desc, a.f_angle, a.f_n, a.f_cap = fnode.data_ptr(), fw.data_ptr(), t['fd'].data_ptr(), t['fa'].data_ptr(), t['fn'].data_ptr(), cap
                a.nnratio, a.check_orientation, a.match_f, a.nmatches = nnratio, ori, m.data_ptr(), nm.data_ptr()
                B.check(B.lib().sgs_match_bow_batch_device(C.byref(a), F, C.c_void_p(0)))
                torch.cuda.synchronize()
                mg, nmg = m.cpu().numpy(), nm.cpu().numpy()
                for f in range(F):
                    a_, b_ = int(kn[f]), int(fn[f])
                    _, okw, okn = V.transform(kd[f, :a_], levelsup); _, ofw, ofn = V.transform(fd[f, :b_], levelsup)
                    onm, om = O.search_by_bow(okn, okw, kv[f, :a_], kd[f, :a_], ka[f, :a_], ofn, ofw, fd[f, :b_], fa[f, :b_], nnratio, bool(ori))
                    assert nmg[f] == onm and np.array_equal(mg[f, :b_], om), (f, levelsup, ori)
                    assert np.all(mg[f, b_:] == -1)
                if flips <= 40 and nk >= 1000:
                    assert nmg[0] > 50
    finally:
        gv.close()


def test_host_variants():
    voc = S.random_vocabulary(21, k=10, L=3)
    V = O.Vocabulary(voc['k'], voc['L'], voc['parent'], voc['desc'], voc['weight'])
    gv = GpuVoc(voc)
    try:
        s = S.bow_pair_scenario(5, voc, n_kf=900, n_f=1000)
        v = C.c_void_p
        outs = []
        for d in (s['kf_desc'], s['f_desc']):
            n = len(d); word = np.zeros(n, np.int32); w = np.zeros(n, np.float64); node = np.zeros(n, np.int32)
            B.check(B.lib().sgs_bow_transform(gv.h, np.ascontiguousarray(d).ctypes.data_as(v), n, 1, word.ctypes.data_as(v), w.ctypes.data_as(v), node.ctypes.data_as(v)))
            ow, owt, on = V.transform(d, 1)
            assert np.array_equal(word, ow) and np.array_equal(node, on) and w.tobytes() == owt.tobytes()
            outs.append((w, node))
        (kw, kn), (fw, fn) = outs
        m = np.zeros(1000, np.int32); nm = C.c_int()
        P = lambda a: np.ascontiguousarray(a).ctypes.data_as(v)
        B.check(B.lib().sgs_match_bow(900, P(kn), P(kw), P(s['kf_valid']), P(s['kf_desc']), P(s['kf_angle']), 1000, P(fn), P(fw), P(s['f_desc']), P(s['f_angle']),
                                      C.c_float(0.7), 1, m.ctypes.data_as(v), C.byref(nm), 0))
        onm, om = O.search_by_bow(kn, kw, s['kf_valid'], s['kf_desc'], s['kf_angle'], fn, fw, s['f_desc'], s['f_angle'], 0.7, True)
        assert nm.value == onm and np.array_equal(m, om)
    finally:
        gv.close()


@pytest.mark.parametrize('seed,n1,n2,nnratio', [(1, 1000, 1000, 0.8), (2, 600, 1100, 0.75)])
def test_search_by_bow_keyframe_pair(seed, n1, n2, nnratio):
    """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (loop closing, src/ORBmatcher.cc:524-657)."""
    import torch
    voc = S.random_vocabulary(13, k=10, L=3)
    V = O.Vocabulary(voc['k'], voc['L'], voc['parent'], voc['desc'], voc['weight'])
    gv = GpuVoc(voc)
    try:
        cap = 1100
        s = S.bow_pair_scenario(seed + 40, voc, n_kf=n1, n_f=n2, flips=40)
        rs = np.random.RandomState(seed)
        valid2 = (rs.rand(n2) < 0.8).astype(np.uint8)
        pad = lambda a, n, shape, dt: np.concatenate([a, np.zeros((cap - n,) + shape, dt)])[None]
        d1 = pad(s['kf_desc'], n1, (32,), np.uint8); d2 = pad(s['f_desc'], n2, (32,), np.uint8)
        a1 = pad(s['kf_angle'], n1, (), np.float32); a2 = pad(s['f_angle'], n2, (), np.float32)
        v1 = pad(s['kf_valid'], n1, (), np.uint8); v2 = pad(valid2, n2, (), np.uint8)
        c1 = np.array([n1], np.int32); c2 = np.array([n2], np.int32)
        _, w1, nd1 = _transform_gpu(gv, d1, c1, 1); _, w2, nd2 = _transform_gpu(gv, d2, c2, 1)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        t = [dev(x) for x in (v1, d1, a1, c1, d2, a2, c2, v2)]
        _, ow1, on1 = V.transform(s['kf_desc'], 1); _, ow2, on2 = V.transform(s['f_desc'], 1)
        for ori in (0, 1):
            m = torch.zeros((1, cap), dtype=torch.int32, device='cuda'); nm = torch.zeros(1, dtype=torch.int32, device='cuda')
            a = B.BowBatch()
            a.kf_node, a.kf_weight, a.kf_valid, a.kf_desc, a.kf_angle, a.kf_n, a.kf_cap = nd1.data_ptr(), w1.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), cap
            a.f_node, a.f_weight, a.f_desc, a.f_angle, a.f_n, a.f_cap = nd2.data_ptr(), w2.data_ptr(), t[4].data_ptr(), t[5].data_ptr(), t[6].data_ptr(), cap
            a.f_valid, a.keyframe_pair = t[7].data_ptr(), 1
            a.nnratio, a.check_orientation, a.match_f, a.nmatches = nnratio, ori, m.data_ptr(), nm.data_ptr()
            B.check(B.lib().sgs_match_bow_batch_device(C.byref(a), 1, C.c_void_p(0)))
            torch.cuda.synchronize()
            onm, om = O.search_by_bow_kfkf(on1, ow1, s['kf_valid'], s['kf_desc'], s['kf_angle'], on2, ow2, valid2, s['f_desc'], s['f_angle'], nnratio, bool(ori))
            assert int(nm.cpu()[0]) == onm and np.array_equal(m.cpu().numpy()[0, :n1], om) and onm > 30
            sel = om >= 0
            assert np.all(valid2[om[sel]] == 1) and len(set(om[sel].tolist())) == int(sel.sum())       # only good map points, each used once
    finally:
        gv.close()
