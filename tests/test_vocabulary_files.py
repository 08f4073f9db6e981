"""Readers of the reference's vocabulary files (ORBVocabulary::loadFromTextFile / loadFromBinaryFile, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1351-1420,
:1467-1508): a random DBoW2-shaped tree is written in both formats exactly as saveToTextFile / saveToBinaryFile lay them out and parsed back into the
flat arrays sgs_vocabulary_create takes.  No device needed (sgs_vocabulary_parse_file is host code)."""
import ctypes as C
import struct

import numpy as np
import pytest

import scenarios as S
from pysgs import binding as B


def write_text(path, voc, trailing_newline=True):
    n = len(voc['parent'])
    child = np.zeros(n, int)
    for i in range(1, n):
        child[voc['parent'][i]] += 1
    with open(path, 'w') as f:
        f.write('%d %d 0 0\n' % (voc['k'], voc['L']))
        for i in range(1, n):
            f.write('%d %d %s %r%s' % (voc['parent'][i], 1 if child[i] == 0 else 0, ' '.join(str(int(b)) for b in voc['desc'][i]), float(voc['weight'][i]),
                                       '\n' if (i < n - 1 or trailing_newline) else ''))
    return child


def write_binary(path, voc):
    n = len(voc['parent'])
    child = np.zeros(n, int)
    for i in range(1, n):
        child[voc['parent'][i]] += 1
    with open(path, 'wb') as f:
        f.write(struct.pack('<IIiiii', n, 41, voc['k'], voc['L'], 0, 0))      # nb_nodes counts the root (saveToBinaryFile, TemplatedVocabulary.h:1517)
        for i in range(1, n):
            f.write(struct.pack('<i', int(voc['parent'][i])) + bytes(voc['desc'][i]) + struct.pack('<f', float(voc['weight'][i])) + bytes([1 if child[i] == 0 else 0]))
    return child


def parse(path):
    lib = B.lib()
    k, L, n = C.c_int(), C.c_int(), C.c_int()
    B.check(lib.sgs_vocabulary_parse_file(str(path).encode(), C.byref(k), C.byref(L), C.byref(n), None, None, None, None, 0))
    parent = np.zeros(n.value, np.int32); desc = np.zeros((n.value, 32), np.uint8); w = np.zeros(n.value, np.float64); leaf = np.zeros(n.value, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.sgs_vocabulary_parse_file(str(path).encode(), None, None, C.byref(n), p(parent), p(desc), p(w), p(leaf), n.value - 1) == B.SGS_ERR_CAPACITY
    B.check(lib.sgs_vocabulary_parse_file(str(path).encode(), None, None, C.byref(n), p(parent), p(desc), p(w), p(leaf), n.value))
    return k.value, L.value, parent, desc, w, leaf


@pytest.mark.parametrize('trailing', [True, False])
def test_text_and_binary_files_round_trip(tmp_path, trailing):
    voc = S.random_vocabulary(5, k=6, L=3)
    n = len(voc['parent'])
    child = write_text(tmp_path / 'voc.txt', voc, trailing)
    k, L, parent, desc, w, leaf = parse(tmp_path / 'voc.txt')
    assert (k, L) == (voc['k'], voc['L']) and len(parent) == n
    assert np.array_equal(parent[1:], voc['parent'][1:]) and np.array_equal(desc[1:], voc['desc'][1:]) and np.array_equal(w[1:], voc['weight'][1:])
    assert np.array_equal(leaf[1:], (child[1:] == 0).astype(np.uint8)) and parent[0] == -1 and not desc[0].any() and w[0] == 0
    write_binary(tmp_path / 'voc.bin', voc)
    k2, L2, parent2, desc2, w2, leaf2 = parse(tmp_path / 'voc.bin')
    assert (k2, L2) == (k, L) and np.array_equal(parent2, parent) and np.array_equal(desc2, desc) and np.array_equal(leaf2, leaf)
    assert np.array_equal(w2[1:], voc['weight'][1:].astype(np.float32).astype(np.float64))          # the binary format stores float weights


def test_bad_files_are_reported(tmp_path):
    lib = B.lib()
    n = C.c_int()
    (tmp_path / 'a.txt').write_text('hello world\n')
    assert lib.sgs_vocabulary_parse_file(str(tmp_path / 'a.txt').encode(), None, None, C.byref(n), None, None, None, None, 0) == B.SGS_ERR_INVALID
    (tmp_path / 'b.txt').write_text('10 6 0 0\n5 1 ' + ' '.join(['1'] * 32) + ' 0.5\n')          # parent 5 does not exist yet
    assert lib.sgs_vocabulary_parse_file(str(tmp_path / 'b.txt').encode(), None, None, C.byref(n), None, None, None, None, 0) == B.SGS_ERR_INVALID
    (tmp_path / 'c.bin').write_bytes(struct.pack('<IIiiii', 3, 41, 10, 6, 0, 0) + b'\\0' * 50)   # truncated
    assert lib.sgs_vocabulary_parse_file(str(tmp_path / 'c.bin').encode(), None, None, C.byref(n), None, None, None, None, 0) == B.SGS_ERR_INVALID
    assert lib.sgs_vocabulary_parse_file(str(tmp_path / 'missing.txt').encode(), None, None, C.byref(n), None, None, None, None, 0) == B.SGS_ERR_INVALID
