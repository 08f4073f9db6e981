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
