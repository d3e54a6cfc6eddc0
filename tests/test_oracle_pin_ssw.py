"""ora_ssw_align (the restated ssw_init + ssw_align, ssw.c:788-941) against the reference's OWN ssw.c,
compiled unmodified into oracle/_ref/libssw_ref.so by oracle/Makefile.ref."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import ora

pytestmark = pytest.mark.skipif(not os.path.exists(ora.REF_SSW), reason="oracle/_ref/libssw_ref.so not built")


class SAlign(C.Structure):  # s_align, include/ssw.h:58-71
    _fields_ = [("cigar", C.POINTER(C.c_uint32)), ("ref_num", C.c_uint32), ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32),
                ("read_begin1", C.c_int32), ("read_end1", C.c_int32), ("readlen", C.c_uint32), ("score1", C.c_uint16),
                ("part", C.c_uint16), ("index_num", C.c_uint16), ("cigarLen", C.c_uint16), ("strand", C.c_bool)]


def ref_ssw(lib, read, ref, mat, go, ge, filters):
    prof = lib.ssw_init(read.ctypes.data_as(C.c_void_p), len(read), mat.ctypes.data_as(C.c_void_p), 5, 2)
    r = lib.ssw_align(prof, ref.ctypes.data_as(C.c_void_p), len(ref), go, ge, 2, filters, 0, 0)
    a = r.contents
    out = (a.score1, a.ref_begin1, a.ref_end1, a.read_begin1, a.read_end1, [a.cigar[i] for i in range(a.cigarLen)])
    pp = C.c_void_p(prof)
    lib.init_destroy(C.byref(pp))
    return out


def make_pair(rng, qlen, tlen, err, nfrac=0.0):
    t = rng.integers(0, 4, tlen).astype(np.int8)
    p = int(rng.integers(0, max(1, tlen - qlen + 1)))
    q = []
    for c in t[p:p + qlen]:
        u = rng.random()
        if u < err / 3:
            continue
        if u < 2 * err / 3:
            q.append(rng.integers(0, 4))
        q.append(rng.integers(0, 4) if rng.random() < err else c)
    q = np.array(q[:qlen] if q else [0], dtype=np.int8)
    if nfrac:
        q[rng.random(q.size) < nfrac] = 4
        t = t.copy(); t[rng.random(t.size) < nfrac] = 4
    flank = rng.integers(0, 4, int(rng.integers(0, 8))).astype(np.int8)
    return np.concatenate([flank, q]), t


@pytest.mark.parametrize("scores", [(2, -3, -3, 5, 2), (2, -4, -2, 6, 3), (1, -3, -3, 5, 2), (2, -7, -7, 3, 1)])
def test_ssw_restatement_matches_reference_ssw(scores):
    lib = C.CDLL(ora.REF_SSW)
    lib.ssw_init.restype = C.c_void_p
    lib.ssw_init.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int8]
    lib.ssw_align.restype = C.POINTER(SAlign)
    lib.ssw_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16, C.c_int32, C.c_int32]
    match, mis, sn, go, ge = scores
    mat = ora.score_matrix(match, mis, sn)
    rng = np.random.default_rng(1234 + match * 7 + go)
    n_word = 0
    for it in range(1500):
        qlen = int(rng.integers(20, 181)) if it % 10 else int(rng.integers(250, 500))
        tlen = qlen + int(rng.integers(-10, 40))
        q, t = make_pair(rng, qlen, max(20, tlen), float(rng.choice([0.0, 0.01, 0.05, 0.15])), nfrac=0.01 if it % 7 == 0 else 0.0)
        filters = int(rng.integers(0, 60))
        exp = ref_ssw(lib, q, t, mat, go, ge, filters)
        rc, out, cig = ora.ssw_align(q, t, mat, go, ge, filters)
        assert rc == 0
        assert int(out[0]) == exp[0] and int(out[2]) == exp[2] and int(out[4]) == exp[4], (it, out, exp[:5])
        if exp[0] >= filters:
            assert (int(out[1]), int(out[3])) == (exp[1], exp[3]), (it, out, exp[:5])
            assert cig.tolist() == exp[5], (it, cig, exp[5])
        n_word += exp[0] >= 255 - abs(min(mis, sn))
    assert n_word > 20  # the int16 (word) path of the reference was exercised too
