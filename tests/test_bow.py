"""Oracle of DBoW2 transform + SearchByBoW: no golden vectors exist (DBoW2 ships no tests, ORBvoc is absent), so the oracle is checked
against an independent numpy restatement of the same definitions on a synthetic vocabulary ("parity unpinned" beyond that)."""
import numpy as np

import oracle as O
import scenarios as S


def _descend_numpy(voc, d, levelsup):
    parent = voc['parent']; n = len(parent)
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[parent[i]].append(i)
    leaf_ids = [i for i in range(1, n) if not children[i]]
    word_of = {nid: w for w, nid in enumerate(leaf_ids)}
    x = np.unpackbits(d)
    node, lvl, nid = 0, 0, 0
    while children[node]:
        lvl += 1
        dist = [int((np.unpackbits(voc['desc'][c]) != x).sum()) for c in children[node]]
        node = children[node][int(np.argmin(dist))]            # argmin returns the first minimum, like the strict '<' of the reference
        if lvl == voc['L'] - levelsup:
            nid = node
    return word_of[node], voc['weight'][node], nid


def test_transform_matches_numpy_restatement():
    voc = S.random_vocabulary(3, k=10, L=3)
    V = O.Vocabulary(voc['k'], voc['L'], voc['parent'], voc['desc'], voc['weight'])
    sc = S.bow_pair_scenario(1, voc, n_kf=300, n_f=10)
    for levelsup in (1, 2, 4):
        word, w, node = V.transform(sc['kf_desc'], levelsup)
        for i in range(0, 300, 7):
            assert (word[i], w[i], node[i]) == _descend_numpy(voc, sc['kf_desc'][i], levelsup)
    word, w, node = V.transform(sc['kf_desc'], 4)
    assert np.all(node == 0)                                   # L - levelsup <= 0: every feature is filed under the root
    ids, vals = O.bow_vector(word, w)
    assert np.all(np.diff(ids) > 0) and abs(vals.sum() - 1.0) < 1e-12


def test_search_by_bow_basic_properties():
    voc = S.random_vocabulary(4, k=10, L=3)
    V = O.Vocabulary(voc['k'], voc['L'], voc['parent'], voc['desc'], voc['weight'])
    sc = S.bow_pair_scenario(2, voc)
    _, wk, nk = V.transform(sc['kf_desc'], 1); _, wf, nf = V.transform(sc['f_desc'], 1)
    for ori in (False, True):
        nm, m = O.search_by_bow(nk, wk, sc['kf_valid'], sc['kf_desc'], sc['kf_angle'], nf, wf, sc['f_desc'], sc['f_angle'], 0.7, ori)
        sel = np.nonzero(m >= 0)[0]
        assert nm == len(sel) and nm > 50
        assert np.all(nk[m[sel]] == nf[sel]) and np.all(sc['kf_valid'][m[sel]] == 1) and np.all(wk[m[sel]] > 0) and np.all(wf[sel] > 0)
        for j in sel[:50]:
            d = int((np.unpackbits(sc['kf_desc'][m[j]]) != np.unpackbits(sc['f_desc'][j])).sum())
            assert d <= 50
    nm0, _ = O.search_by_bow(nk, wk, sc['kf_valid'], sc['kf_desc'], sc['kf_angle'], nf, wf, sc['f_desc'], sc['f_angle'], 0.7, False)
    nm1, _ = O.search_by_bow(nk, wk, sc['kf_valid'], sc['kf_desc'], sc['kf_angle'], nf, wf, sc['f_desc'], sc['f_angle'], 0.7, True)
    assert nm1 <= nm0


def _triangulation_python(k1, k2, F12, ex, ey, sigma2, scale, only_stereo):
    """Plain-Python transcription of the loop at src/ORBmatcher.cc:689-773 (without the orientation filter) on numpy float32 scalars."""
    f32 = np.float32
    fv1, fv2 = {}, {}
    for i in range(len(k1['node'])):
        if k1['weight'][i] > 0:
            fv1.setdefault(int(k1['node'][i]), []).append(i)
    for j in range(len(k2['node'])):
        if k2['weight'][j] > 0:
            fv2.setdefault(int(k2['node'][j]), []).append(j)
    m = np.full(len(k1['node']), -1, np.int32); matched2 = np.zeros(len(k2['node']), bool)
    Fm = np.asarray(F12, f32)
    for node in sorted(set(fv1) & set(fv2)):
        for i1 in fv1[node]:
            if not k1['free'][i1] or (only_stereo and not k1['stereo'][i1]):
                continue
            x1, y1 = f32(k1['xy'][i1, 0]), f32(k1['xy'][i1, 1])
            best, bi = 50, -1
            for i2 in fv2[node]:
                if matched2[i2] or not k2['free'][i2] or (only_stereo and not k2['stereo'][i2]):
                    continue
                d = int((np.unpackbits(k1['desc'][i1]) != np.unpackbits(k2['desc'][i2])).sum())
                if d > 50 or d > best:
                    continue
                x2, y2 = f32(k2['xy'][i2, 0]), f32(k2['xy'][i2, 1]); oc = int(k2['octave'][i2])
                if not k1['stereo'][i1] and not k2['stereo'][i2]:
                    dx, dy = f32(ex) - x2, f32(ey) - y2
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(100 * scale[oc]):
                        continue
                a = f32(f32(f32(x1 * Fm[0, 0]) + f32(y1 * Fm[1, 0])) + Fm[2, 0]); b = f32(f32(f32(x1 * Fm[0, 1]) + f32(y1 * Fm[1, 1])) + Fm[2, 1])
                c = f32(f32(f32(x1 * Fm[0, 2]) + f32(y1 * Fm[1, 2])) + Fm[2, 2])
                num = f32(f32(f32(a * x2) + f32(b * y2)) + c); den = f32(f32(a * a) + f32(b * b))
                if den == 0:
                    continue
                if float(f32(f32(num * num) / den)) < 3.84 * float(sigma2[oc]):
                    best, bi = d, i2
            if bi >= 0:
                m[i1] = bi; matched2[bi] = True
    return int((m >= 0).sum()), m


def test_search_for_triangulation_matches_python_transcription():
    voc = S.random_vocabulary(6, k=6, L=2)
    V = O.Vocabulary(voc['k'], voc['L'], voc['parent'], voc['desc'], voc['weight'])
    rs = np.random.RandomState(2)
    n1, n2 = 220, 260
    s = S.bow_pair_scenario(11, voc, n_kf=n1, n_f=n2, flips=25)
    d1b = np.unpackbits(s['kf_desc'], axis=1).astype(np.int16); d2b = np.unpackbits(s['f_desc'], axis=1).astype(np.int16)
    src = np.array([int(np.argmin(np.abs(d1b - d2b[j]).sum(1))) for j in range(n2)])
    xy1 = np.c_[rs.uniform(20, 620, n1), rs.uniform(20, 460, n1)].astype(np.float32)
    xy2 = np.c_[xy1[src, 0] + rs.uniform(-40, 40, n2), xy1[src, 1] + rs.normal(0, 1.5, n2)].astype(np.float32)
    sf = S.scale_factors().astype(np.float32); sigma2 = (sf * sf).astype(np.float32)
    _, w1, nd1 = V.transform(s['kf_desc'], 1); _, w2, nd2 = V.transform(s['f_desc'], 1)
    k1 = dict(node=nd1, weight=w1, free=(rs.rand(n1) < 0.8).astype(np.uint8), stereo=(rs.rand(n1) < 0.5).astype(np.uint8), desc=s['kf_desc'], xy=xy1, angle=s['kf_angle'])
    k2 = dict(node=nd2, weight=w2, free=(rs.rand(n2) < 0.8).astype(np.uint8), stereo=(rs.rand(n2) < 0.5).astype(np.uint8), desc=s['f_desc'], xy=xy2,
              octave=rs.randint(0, 8, n2).astype(np.int32), angle=s['f_angle'])
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    for only in (False, True):
        nm, m = O.search_for_triangulation(k1, k2, F12, 320.5, 240.25, sigma2, sf, only, False)
        pn, pm = _triangulation_python(k1, k2, F12, 320.5, 240.25, sigma2, sf, only)
        assert nm == pn and np.array_equal(m, pm) and nm > 5
