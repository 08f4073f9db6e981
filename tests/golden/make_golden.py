#!/usr/bin/env python3
"""Generate the golden vectors that PIN the CPU oracle (tests/golden/*.npz).

This is an INDEPENDENT Python restatement of the reference extractor
(/root/reference/src/sg-slam/src/ORBextractor.cc) in which every OpenCV call the reference makes is
made through the REAL OpenCV (cv2): cv2.resize (ORBextractor.cc:1121), cv2.FastFeatureDetector
(:810,:815), cv2.GaussianBlur (:1087), cv2.fastAtan2 (:104).  Only the ORB-SLAM-specific glue
(cell loop, quadtree, moment loops, rotated-BRIEF taps) is restated by hand, separately from
oracle/sgs_oracle.cpp.  The C++ oracle (which uses no OpenCV) must reproduce these vectors
bit-for-bit (tests/test_oracle_golden.py); the reference itself ships no tests or fixtures
(SURVEY.md section 4), so this is the strongest pin available.

Run in the build container (needs cv2):  python tests/golden/make_golden.py
"""
import math
import os
import re
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'sg-slam_b200'))
from pysgs import synth  # noqa: E402

f32 = np.float32


def load_pattern():
    txt = open(os.path.join(ROOT, 'oracle', 'orb_pattern.inc')).read()
    txt = '\n'.join(l for l in txt.splitlines() if not l.startswith('//'))
    v = np.array([int(x) for x in re.findall(r'-?\d+', txt)], np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)


PATTERN = load_pattern()
EDGE = 19
HALF = 15


def cv_round(x):
    return int(np.rint(np.float64(x)))


def tables(nfeatures, scaleFactor, nlevels):
    sf = np.float64(f32(scaleFactor))
    scale = [f32(1.0)]
    for i in range(1, nlevels):
        scale.append(f32(np.float64(scale[-1]) * sf))
    inv = [f32(1.0) / s for s in scale]
    factor = f32(1.0 / sf)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    nd = f32(nd)
    per = []
    for _ in range(nlevels - 1):
        per.append(cv_round(nd))
        nd = f32(nd * factor)
    per.append(max(nfeatures - sum(per), 0))
    umax = [0] * 16
    vmax = int(math.floor(float(f32(HALF) * f32(math.sqrt(2.0)) / f32(2) + f32(1))))
    vmin = int(math.ceil(float(f32(HALF) * f32(math.sqrt(2.0)) / f32(2))))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(HALF * HALF - v * v))
    v0 = 0
    for v in range(HALF, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return scale, inv, per, umax


class Node:
    __slots__ = ('keys', 'UL', 'UR', 'BL', 'BR', 'noMore', 'seq', 'alive')

    def __init__(self):
        self.keys = []
        self.noMore = False
        self.alive = True


def divide(n, c):
    halfX = int(math.ceil(float(f32(n.UR[0] - n.UL[0]) / f32(2))))
    halfY = int(math.ceil(float(f32(n.BR[1] - n.UL[1]) / f32(2))))
    n1, n2, n3, n4 = Node(), Node(), Node(), Node()
    n1.UL = n.UL; n1.UR = (n.UL[0] + halfX, n.UL[1]); n1.BL = (n.UL[0], n.UL[1] + halfY); n1.BR = (n.UL[0] + halfX, n.UL[1] + halfY)
    n2.UL = n1.UR; n2.UR = n.UR; n2.BL = n1.BR; n2.BR = (n.UR[0], n.UL[1] + halfY)
    n3.UL = n1.BL; n3.UR = n1.BR; n3.BL = n.BL; n3.BR = (n1.BR[0], n.BL[1])
    n4.UL = n3.UR; n4.UR = n2.BR; n4.BL = n3.BR; n4.BR = n.BR
    for k in n.keys:
        x, y = c[k][0], c[k][1]
        if x < n1.UR[0]:
            (n1 if y < n1.BR[1] else n3).keys.append(k)
        elif y < n1.BR[1]:
            n2.keys.append(k)
        else:
            n4.keys.append(k)
    for ch in (n1, n2, n3, n4):
        ch.noMore = len(ch.keys) == 1
    return n1, n2, n3, n4


def octree(c, minX, maxX, minY, maxY, N):
    """DistributeOctTree (ORBextractor.cc:540-764) on a python list acting as std::list (front = index 0).
    Tie-break of the (size, pointer) sort := node creation sequence (quirk Q1)."""
    if not len(c):
        return []
    nIni = int(np.round(f32(maxX - minX) / f32(maxY - minY)))   # std::round(float): half away from zero
    r = float(f32(maxX - minX) / f32(maxY - minY))
    nIni = int(math.floor(r + 0.5)) if r >= 0 else -int(math.floor(-r + 0.5))
    hX = f32(maxX - minX) / f32(nIni)
    nodes = []
    seq = [0]
    ini = []
    for i in range(nIni):
        n = Node()
        n.UL = (int(hX * f32(i)), 0); n.UR = (int(hX * f32(i + 1)), 0)
        n.BL = (n.UL[0], maxY - minY); n.BR = (n.UR[0], maxY - minY)
        n.seq = seq[0]; seq[0] += 1
        nodes.append(n); ini.append(n)
    for i in range(len(c)):
        ini[int(f32(c[i][0]) / hX)].keys.append(i)
    nodes = [n for n in nodes if n.keys]
    for n in nodes:
        if len(n.keys) == 1:
            n.noMore = True
    finish = False
    while not finish:
        prevSize = len(nodes)
        nToExpand = 0
        sz = []
        # coarse pass: children are pushed to the FRONT, the walk only moves forward
        front = []
        rest = []
        for n in nodes:
            if n.noMore:
                rest.append(n)
                continue
            for ch in divide(n, c):
                if ch.keys:
                    ch.seq = seq[0]; seq[0] += 1
                    front.insert(0, ch)
                    if len(ch.keys) > 1:
                        nToExpand += 1
                        sz.append(ch)
        nodes = front + rest
        if len(nodes) >= N or len(nodes) == prevSize:
            finish = True
        elif len(nodes) + nToExpand * 3 > N:
            while not finish:
                prevSize = len(nodes)
                prev = sorted(sz, key=lambda n: (len(n.keys), n.seq))
                sz = []
                for n in reversed(prev):
                    for ch in divide(n, c):
                        if ch.keys:
                            ch.seq = seq[0]; seq[0] += 1
                            nodes.insert(0, ch)
                            if len(ch.keys) > 1:
                                sz.append(ch)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prevSize:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if c[k][2] > c[best][2]:
                best = k
        out.append(best)
    return out


def extract(img, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniTh=20, minTh=7):
    scale, inv, per, umax = tables(nfeatures, scaleFactor, nlevels)
    h, w = img.shape
    pyr = []
    for l in range(nlevels):
        lw = cv_round(f32(w) * inv[l]); lh = cv_round(f32(h) * inv[l])
        if l == 0:
            pyr.append(img.copy())
        else:
            pyr.append(cv2.resize(pyr[l - 1], (lw, lh), interpolation=cv2.INTER_LINEAR))
    det_ini = cv2.FastFeatureDetector_create(iniTh, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(minTh, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    all_kps = []
    all_cands = []
    for l in range(nlevels):
        L = pyr[l]
        minBX = minBY = EDGE - 3
        maxBX = L.shape[1] - EDGE + 3; maxBY = L.shape[0] - EDGE + 3
        width = f32(maxBX - minBX); height = f32(maxBY - minBY)
        nCols = int(width / f32(30)); nRows = int(height / f32(30))
        wCell = int(math.ceil(float(width / f32(nCols)))); hCell = int(math.ceil(float(height / f32(nRows))))
        cands = []
        for i in range(nRows):
            iniY = minBY + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                view = np.ascontiguousarray(L[iniY:maxY, iniX:maxX])
                k = det_ini.detect(view)
                if not k:
                    k = det_min.detect(view)
                for p in k:
                    cands.append((f32(p.pt[0] + j * wCell), f32(p.pt[1] + i * hCell), f32(p.response)))
        sel = octree(cands, minBX, maxBX, minBY, maxBY, per[l])
        kps = []
        size = f32(int(f32(31) * scale[l]))
        for s in sel:
            x = f32(cands[s][0] + minBX); y = f32(cands[s][1] + minBY)
            kps.append([x, y, size, f32(-1), cands[s][2], l])
        all_kps.append(kps)
        all_cands.append(np.array(cands, np.float32).reshape(-1, 3))
    # orientation (IC_Angle :78-105)
    for l in range(nlevels):
        L = pyr[l].astype(np.int64)
        for k in all_kps[l]:
            cx, cy = cv_round(k[0]), cv_round(k[1])
            m01 = 0; m10 = 0
            for u in range(-HALF, HALF + 1):
                m10 += u * int(L[cy, cx + u])
            for v in range(1, HALF + 1):
                d = umax[v]
                us = np.arange(-d, d + 1)
                plus = L[cy + v, cx - d:cx + d + 1]; minus = L[cy - v, cx - d:cx + d + 1]
                m10 += int((us * (plus + minus)).sum())
                m01 += v * int((plus - minus).sum())
            k[3] = f32(cv2.fastAtan2(float(m01), float(m10)))
    out_kps = []
    out_desc = []
    factorPI = f32(np.float64(np.pi) / np.float64(f32(180.0)))
    px = PATTERN[:, 0].astype(np.float32); py = PATTERN[:, 1].astype(np.float32)
    blurred_all = []
    for l in range(nlevels):
        if not all_kps[l]:
            blurred_all.append(None)
            continue
        B = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        blurred_all.append(B)
        for k in all_kps[l]:
            ang = f32(k[3]) * factorPI
            a = f32(np.cos(np.float64(ang))); b = f32(np.sin(np.float64(ang)))
            cx, cy = cv_round(k[0]), cv_round(k[1])
            ry = np.rint((px * b + py * a).astype(np.float32).astype(np.float64)).astype(np.int64)
            rx = np.rint((px * a - py * b).astype(np.float32).astype(np.float64)).astype(np.int64)
            vals = B[cy + ry, cx + rx].astype(np.int32)
            bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
            out_desc.append(np.packbits(bits, bitorder='little'))
            out_kps.append((f32(k[0] * scale[l]) if l else k[0], f32(k[1] * scale[l]) if l else k[1], k[2], k[3], k[4], l, -1))
    kp_dtype = np.dtype([('x', '<f4'), ('y', '<f4'), ('size', '<f4'), ('angle', '<f4'), ('response', '<f4'), ('octave', '<i4'), ('class_id', '<i4')])
    kps = np.array(out_kps, kp_dtype)
    desc = np.array(out_desc, np.uint8).reshape(-1, 32)
    return dict(kps=kps, desc=desc, pyramid=pyr, blurred=blurred_all, cands=all_cands, per=per, umax=umax, scale=np.array(scale, np.float32))


def main():
    cases = {
        's1_640x480': (synth.frame_s1(640, 480, 1), dict(nfeatures=1000)),
        's1_320x240': (synth.frame_s1(320, 240, 3), dict(nfeatures=500)),
        'noise_200x160': (np.random.RandomState(7).randint(0, 256, (160, 200)).astype(np.uint8), dict(nfeatures=300, nlevels=4)),
    }
    for name, (img, kw) in cases.items():
        r = extract(img, **kw)
        lv_sizes = np.array([[p.shape[1], p.shape[0]] for p in r['pyramid']], np.int32)
        ncands = np.array([len(c) for c in r['cands']], np.int32)
        # checksums of per-level images instead of the images themselves (keeps fixtures small)
        pyr_sums = np.array([int(p.astype(np.uint64).sum()) for p in r['pyramid']], np.uint64)
        pyr_xor = np.array([int(np.bitwise_xor.reduce(p.reshape(-1).astype(np.uint64) * (np.arange(p.size, dtype=np.uint64) % 65521 + 1))) for p in r['pyramid']], np.uint64)
        blur_sums = np.array([int(b.astype(np.uint64).sum()) if b is not None else 0 for b in r['blurred']], np.uint64)
        last = len(r['pyramid']) - 1
        np.savez_compressed(os.path.join(HERE, 'extract_%s.npz' % name), image=img, kps=r['kps'], desc=r['desc'],
                            level_sizes=lv_sizes, ncands=ncands, pyr_sums=pyr_sums, pyr_xor=pyr_xor, blur_sums=blur_sums,
                            cands_l0=r['cands'][0], cands_last=r['cands'][last], pyr_last=r['pyramid'][last],
                            blur_last=r['blurred'][last] if r['blurred'][last] is not None else np.zeros((0, 0), np.uint8),
                            per=np.array(r['per'], np.int32), umax=np.array(r['umax'], np.int32), scale=r['scale'],
                            params=np.array([kw.get('nfeatures', 1000), kw.get('nlevels', 8), 20, 7], np.int32),
                            cv2_version=np.array(cv2.__version__))
        print(name, 'kps', len(r['kps']), 'cands', ncands.tolist(), 'per', r['per'])
    # primitive-level vectors: fastAtan2 and small resize/blur cases
    rng = np.random.RandomState(11)
    ys = rng.randint(-200000, 200000, 4000).astype(np.float32); xs = rng.randint(-200000, 200000, 4000).astype(np.float32)
    ys[:8] = [0, 0, 1, -1, 5, -5, 0, 3]; xs[:8] = [0, 1, 0, 0, 5, 5, -2, -3]
    at = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    small = rng.randint(0, 256, (37, 53)).astype(np.uint8)
    rs = cv2.resize(small, (44, 31), interpolation=cv2.INTER_LINEAR)
    bl = cv2.GaussianBlur(small, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    fd = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fk = fd.detect(small)
    fast_kps = np.array([[p.pt[0], p.pt[1], p.response] for p in fk], np.float32).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, 'primitives.npz'), atan_y=ys, atan_x=xs, atan=at, small=small, resize_44x31=rs,
                        blur=bl, fast20=fast_kps, cv2_version=np.array(cv2.__version__))
    print('primitives ok, fast kps', len(fast_kps))


if __name__ == '__main__':
    main()
