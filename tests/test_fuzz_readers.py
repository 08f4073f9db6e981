"""The file readers behind the C ABI -- sgs_vocabulary_parse_file (DBoW2 text / binary), the detector's ncnn .param / .bin reader (sgs_detector_create in plan-only
mode) and sgs_settings_load -- on mutated inputs: random byte edits, truncations, insertions, hostile header fields and counts, deleted / swapped graph lines.
Every call must come back with a status code (OK, INVALID, UNSUPPORTED, CAPACITY); nothing may crash, hang or throw through the C boundary, and a file that is
accepted must describe a tree the device loader would accept too (parents name earlier nodes).  Host code only: no device needed.  Seeds are fixed."""
import ctypes as C
import os
import re
import struct

import numpy as np

import detector_model as DM
import scenarios as S
import test_settings as TS
import test_vocabulary_files as TV
from pysgs import binding as B

OK_CODES = {B.SGS_OK, B.SGS_ERR_INVALID, B.SGS_ERR_UNSUPPORTED, B.SGS_ERR_CAPACITY}


def _mutate(rng, good, binary_header=None):
    b = bytearray(good)
    mode = rng.randint(4)
    if mode == 0:
        for _ in range(rng.randint(1, 6)):
            b[rng.randint(len(b))] = rng.randint(256)
    elif mode == 1:
        b = b[:rng.randint(len(b))]
    elif mode == 2:
        i = rng.randint(len(b)); b[i:i] = bytes(rng.randint(0, 256, rng.randint(1, 40)).astype(np.uint8))
    elif binary_header:
        hdr = list(struct.unpack(binary_header, bytes(b[:struct.calcsize(binary_header)])))
        hdr[rng.randint(len(hdr))] = int(rng.choice([0, 1, 2, 12345678, 2 ** 31 - 1]))
        b[:struct.calcsize(binary_header)] = struct.pack(binary_header, *hdr)
    else:
        i = rng.randint(len(b)); b[i:i + rng.randint(1, 8)] = b'-999999999999'
    return bytes(b)


def test_vocabulary_readers_survive_mutated_files(tmp_path):
    lib = B.lib()
    voc = S.random_vocabulary(5, k=4, L=2)
    TV.write_text(tmp_path / 'v.txt', voc); TV.write_binary(tmp_path / 'v.bin', voc)
    rng = np.random.RandomState(0)
    seen = set()
    for ext, hdr in (('.txt', None), ('.bin', '<IIiiii')):
        good = open(tmp_path / ('v' + ext), 'rb').read()
        for _ in range(500):
            p = str(tmp_path / ('f' + ext))
            open(p, 'wb').write(_mutate(rng, good, hdr))
            k, L, n = C.c_int(), C.c_int(), C.c_int()
            rc = lib.sgs_vocabulary_parse_file(p.encode(), C.byref(k), C.byref(L), C.byref(n), None, None, None, None, 0)
            assert rc in OK_CODES
            if rc == B.SGS_OK:
                assert 2 <= n.value < 10_000_000
                parent = np.zeros(n.value, np.int32); desc = np.zeros((n.value, 32), np.uint8); w = np.zeros(n.value, np.float64); leaf = np.zeros(n.value, np.uint8)
                q = lambda a: a.ctypes.data_as(C.c_void_p)
                assert lib.sgs_vocabulary_parse_file(p.encode(), None, None, C.byref(n), q(parent), q(desc), q(w), q(leaf), n.value) == B.SGS_OK
                assert parent[0] == -1 and (parent[1:] >= 0).all() and (parent[1:] < np.arange(1, n.value)).all()
            seen.add(rc)
    assert B.SGS_OK in seen and B.SGS_ERR_INVALID in seen


def test_detector_graph_reader_survives_mutated_files(tmp_path):
    pp, bp = DM.write_mini_model(str(tmp_path), 0)
    good = open(pp).read()
    toks = re.findall(r'\S+|\s+', good)
    rng = np.random.RandomState(1)
    seen = {}
    for _ in range(400):
        mode = rng.randint(5)
        if mode == 0:
            t = list(toks)
            for _ in range(rng.randint(1, 4)):
                i = rng.randint(len(t))
                if not t[i].isspace():
                    t[i] = str(rng.choice(['-1', '0', '99999999', '-233', '2147483647', 'abc', '1.5e38', '', '-23300=9999999,1', '3=-5', '0=0']))
            txt = ''.join(t)
        elif mode == 1:
            txt = good[:rng.randint(len(good))]
        elif mode == 2:
            lines = good.split('\n'); del lines[rng.randint(len(lines))]; txt = '\n'.join(lines)
        elif mode == 3:
            lines = good.split('\n'); i, j = rng.randint(2, len(lines)), rng.randint(2, len(lines)); lines[i], lines[j] = lines[j], lines[i]; txt = '\n'.join(lines)
        else:
            txt = _mutate(rng, good.encode()).decode('latin1')
        p = str(tmp_path / 'f.param')
        open(p, 'w', encoding='latin1').write(txt)
        try:
            det = B.Detector(p, bp, flags=B.DET_PLAN_ONLY); rc = B.SGS_OK
            det.close()
        except B.SgsError as e:
            rc = e.code
        assert rc in OK_CODES
        seen[rc] = seen.get(rc, 0) + 1
    assert seen.get(B.SGS_ERR_INVALID, 0) > 100 and seen.get(B.SGS_OK, 0) > 5
    # the weight file: truncated and grown
    blob = open(bp, 'rb').read()
    for cut in (0, 1, 3, len(blob) // 2, len(blob) - 1):
        q = str(tmp_path / 'w.bin'); open(q, 'wb').write(blob[:cut])
        try:
            B.Detector(pp, q, flags=B.DET_PLAN_ONLY).close(); rc = B.SGS_OK
        except B.SgsError as e:
            rc = e.code
        assert rc == B.SGS_ERR_INVALID


def test_settings_reader_survives_mutated_files(tmp_path):
    lib = B.lib()
    rng = np.random.RandomState(2)
    good = TS.YAML.encode()
    for _ in range(400):
        p = str(tmp_path / 's.yaml')
        open(p, 'wb').write(_mutate(rng, good))
        s = B.Settings()
        assert lib.sgs_settings_load(p.encode(), C.byref(s)) in OK_CODES
