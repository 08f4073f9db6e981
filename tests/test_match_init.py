"""CPU check of the oracle's restatement of ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:407-522) against an independent numpy
transcription of the loop (candidates from the separately tested GetFeaturesInArea restatement) on scenarios where several reference keypoints
compete for the same current-frame feature, so that the distance book-keeping and the stealing of matches are exercised."""
import numpy as np
import pytest

import oracle as O
import scenarios as S

POP = np.array([bin(i).count('1') for i in range(256)], np.int32)


def init_scenario(seed, n1=800, n2=900):
    rs = np.random.RandomState(seed)
    cam = dict(fx=535.4, fy=539.2, cx=320.1, cy=247.6, bf=40.0)
    base = rs.randint(0, 256, (n1, 32)).astype(np.uint8)
    xy1 = np.c_[rs.uniform(5, 635, n1), rs.uniform(5, 475, n1)].astype(np.float32)
    dup = rs.rand(n1) < 0.25                                   # near-duplicates of an earlier keypoint: rivals for the same feature
    for i in np.nonzero(dup)[0]:
        if i == 0:
            continue
        j = rs.randint(0, i)
        b = np.unpackbits(base[j]); b[rs.choice(256, rs.randint(0, 6), replace=False)] ^= 1
        base[i] = np.packbits(b); xy1[i] = xy1[j] + rs.normal(0, 8, 2)
    k1 = np.zeros(n1, O.KP_DTYPE); k1['x'] = xy1[:, 0]; k1['y'] = xy1[:, 1]
    k1['octave'] = np.where(rs.rand(n1) < 0.6, 0, rs.randint(1, 8, n1)); k1['angle'] = rs.uniform(0, 360, n1)
    src = rs.randint(0, n1, n2)
    bits = np.unpackbits(base[src], axis=1)
    for j in range(n2):
        bits[j, rs.choice(256, rs.randint(0, 50), replace=False)] ^= 1
    d2 = np.packbits(bits, axis=1)
    k2 = np.zeros(n2, O.KP_DTYPE)
    k2['x'] = np.clip(xy1[src, 0] + rs.normal(0, 25, n2), 1, 638); k2['y'] = np.clip(xy1[src, 1] + rs.normal(0, 25, n2), 1, 478)
    k2['octave'] = np.where(rs.rand(n2) < 0.7, 0, rs.randint(1, 8, n2))
    k2['angle'] = (k1['angle'][src] - 20 + rs.normal(0, 5, n2) + (rs.rand(n2) < 0.2) * rs.uniform(0, 360, n2)) % 360
    sf = S.scale_factors()
    f1 = O.FrameArrays(k1, np.full(n1, -1, np.float32), base, 640, 480, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf'], sf)
    f2 = O.FrameArrays(k2, np.full(n2, -1, np.float32), d2, 640, 480, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf'], sf)
    prev = np.stack([k1['x'], k1['y']], 1).astype(np.float32)
    return dict(k1=k1, d1=base, k2=k2, d2=d2, f1=f1, f2=f2, prev=prev, cam=cam, sf=sf)


def transcription(s, window, nnratio, check_ori):
    k1, k2, d1, d2 = s['k1'], s['k2'], s['d1'], s['d2']
    n1, n2 = len(k1), len(k2)
    m12 = np.full(n1, -1, np.int32); m21 = np.full(n2, -1, np.int32); md = np.full(n2, np.iinfo(np.int32).max, np.int64)
    hist = [[] for _ in range(30)]
    nm = 0
    for i1 in range(n1):
        if k1['octave'][i1] > 0:
            continue
        cand = O.features_in_area(s['f2'], float(s['prev'][i1, 0]), float(s['prev'][i1, 1]), float(window), 0, 0)
        best, best2, bi = 2 ** 31 - 1, 2 ** 31 - 1, -1
        for i2 in cand:
            d = int(POP[d1[i1] ^ d2[i2]].sum())
            if md[i2] <= d:
                continue
            if d < best:
                best2, best, bi = best, d, i2
            elif d < best2:
                best2 = d
        if best <= 50 and np.float32(best) < np.float32(np.float32(best2) * np.float32(nnratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1; nm -= 1
            m12[i1] = bi; m21[bi] = i1; md[bi] = best; nm += 1
            if check_ori:
                rot = np.float32(k1['angle'][i1] - k2['angle'][bi])
                if rot < 0:
                    rot = np.float32(rot + np.float32(360))
                v = float(np.float32(rot * np.float32(30 / 360.0)))
                b = int(np.floor(v + 0.5)) if v >= 0 else int(np.ceil(v - 0.5))
                hist[0 if b == 30 else b].append(i1)
    if check_ori:
        sizes = [len(h) for h in hist]
        order = sorted(range(30), key=lambda i: (-sizes[i], i))
        m1, m2_, m3 = sizes[order[0]], sizes[order[1]], sizes[order[2]]
        keep = [order[0]] + ([order[1]] if m2_ >= 0.1 * m1 else []) + ([order[2]] if m2_ >= 0.1 * m1 and m3 >= 0.1 * m1 else [])
        for i in range(30):
            if i in keep:
                continue
            for idx in hist[i]:
                if m12[idx] >= 0:
                    m12[idx] = -1; nm -= 1
    prev = s['prev'].copy()
    for i1 in np.nonzero(m12 >= 0)[0]:
        prev[i1] = (k2['x'][m12[i1]], k2['y'][m12[i1]])
    return nm, m12, prev


@pytest.mark.parametrize('seed,window,ori', [(1, 100, True), (2, 100, False), (3, 40, True)])
def test_oracle_matches_the_transcription(seed, window, ori):
    s = init_scenario(seed)
    nm, m, prev = O.search_for_initialization(s['f1'], s['f2'], s['prev'], window, 0.9, ori)
    n2, m2, p2 = transcription(s, window, 0.9, ori)
    assert nm == n2 and np.array_equal(m, m2) and np.array_equal(prev, p2)
    assert nm > 40 and (m >= 0).sum() == nm and len(np.unique(m[m >= 0])) == nm          # one-to-one
    assert np.all(s['k1']['octave'][m >= 0] == 0) and np.all(s['k2']['octave'][m[m >= 0]] == 0)
