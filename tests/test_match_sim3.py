"""CPU check of the oracle's restatement of ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (src/ORBmatcher.cc:292-405)
against an independent numpy transcription of the same loop (float32 arithmetic rounded per operation, candidates from the separately tested
GetFeaturesInArea restatement): identical vpMatched, identical count, on scenarios where many points compete for the same features."""
import numpy as np
import pytest

import oracle as O
import scenarios as S

POP = np.array([bin(i).count('1') for i in range(256)], np.int32)


def sim3_inputs(seed, ncur, nmp):
    s = S.keyframe_scenario(seed, n_cur=ncur, n_kf=nmp, conflict=0.4)
    rs = np.random.RandomState(seed + 7)
    R = s['Tcw_cur'][:3, :3].astype(np.float64); t = s['Tcw_cur'][:3, 3].astype(np.float64)
    Ow = (-(R.T @ t)).astype(np.float32)
    to = s['last_xyz'].astype(np.float64) - Ow.astype(np.float64); d = np.linalg.norm(to, axis=1)
    nrm = to / np.maximum(d[:, None], 1e-9) + rs.normal(0, 0.6, (nmp, 3)); nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    matched = np.where(rs.rand(ncur) < 0.1, 100000 + np.arange(ncur), -1).astype(np.int32)      # vpMatched entries that are already taken
    return s, Ow, nrm, matched


def transcription(s, Ow, nrm, matched, th):
    f32 = np.float32
    cam = s['cam']; sf = s['sf'].astype(f32)
    fo = O.FrameArrays(s['kps'], s['uright'], s['desc'], 640, 480, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf'], s['sf'])
    T = s['Tcw_cur'].astype(f32); R = T[:3, :3]; t = T[:3, 3]
    fx, fy, cx, cy = [f32(cam[k]) for k in ('fx', 'fy', 'cx', 'cy')]
    logsf = O.logf(1.2)
    m = matched.copy(); n = 0
    for i in range(len(s['last_xyz'])):
        if not s['kf_valid'][i]:
            continue
        X = s['last_xyz'][i].astype(f32)
        pc = [f32(np.float64(f32(f32(R[r, 0] * X[0]) + f32(R[r, 1] * X[1])) + f32(R[r, 2] * X[2])) + np.float64(t[r])) for r in range(3)]
        if pc[2] < 0:
            continue
        invz = f32(1) / pc[2]
        u = f32(fx * f32(pc[0] * invz)) + cx; v = f32(fy * f32(pc[1] * invz)) + cy
        if not (0 <= u < 640 and 0 <= v < 480):
            continue
        PO = X - Ow
        dist = f32(np.sqrt(np.float64(PO[0]) * PO[0] + np.float64(PO[1]) * PO[1] + np.float64(PO[2]) * PO[2]))
        if dist < f32(0.8) * s['min_dist'][i] or dist > f32(1.2) * s['max_dist'][i]:
            continue
        if (np.float64(PO[0]) * nrm[i, 0] + np.float64(PO[1]) * nrm[i, 1]) + np.float64(PO[2]) * nrm[i, 2] < 0.5 * np.float64(dist):
            continue
        lvl = int(np.ceil(f32(O.logf(f32(s['max_dist'][i] / dist)) / logsf)))
        lvl = min(max(lvl, 0), 7)
        cand = O.features_in_area(fo, float(u), float(v), float(f32(th) * sf[lvl]))
        best, bi = 256, -1
        for idx in cand:
            if m[idx] >= 0 or not (lvl - 1 <= s['kps']['octave'][idx] <= lvl):
                continue
            d = int(POP[s['last_desc'][i] ^ s['desc'][idx]].sum())
            if d < best:
                best, bi = d, idx
        if best <= 50:
            m[bi] = i; n += 1
    return n, m


@pytest.mark.parametrize('seed,ncur,nmp,th', [(11, 600, 900, 10), (12, 300, 1500, 10), (13, 800, 800, 4)])
def test_oracle_matches_the_transcription(seed, ncur, nmp, th):
    s, Ow, nrm, matched = sim3_inputs(seed, ncur, nmp)
    cam = s['cam']
    fo = O.FrameArrays(s['kps'], s['uright'], s['desc'], 640, 480, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf'], s['sf'])
    nm, m = O.search_by_projection_sim3(fo, s['Tcw_cur'], Ow, s['kf_valid'], s['last_xyz'], nrm, s['min_dist'], s['max_dist'], s['last_desc'], float(th), matched)
    n2, m2 = transcription(s, Ow, nrm, matched, th)
    assert nm == n2 and np.array_equal(m, m2)
    assert nm > 30 and np.array_equal(m[matched >= 0], matched[matched >= 0])          # occupied entries are never overwritten
    # order dependence is real in these scenarios: visiting the points backwards gives another assignment
    rev = slice(None, None, -1)
    nm_r, m_r = O.search_by_projection_sim3(fo, s['Tcw_cur'], Ow, s['kf_valid'][rev], s['last_xyz'][rev], nrm[rev], s['min_dist'][rev], s['max_dist'][rev],
                                            s['last_desc'][rev], float(th), matched)
    new = (matched < 0) & (m >= 0)
    assert not np.array_equal(np.where(new, len(nrm) - 1 - m_r, -1)[new], m[new])
