"""ctypes wrapper of the CPU oracle (oracle/libsmr_oracle.so) and of the reference build under
oracle/_ref/.  TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and the CPU
legs of bench.py -- never by sortmerna_b200/."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsmr_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "sortmerna_ref")
REF_SSW = os.path.join(HERE, "_ref", "libssw_ref.so")


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "match", "mismatch", "score_N", "gap_open", "gap_ext", "num_seeds", "min_lis", "edges",
        "edges_is_percent", "num_alignments", "is_best", "is_forward", "is_reverse", "is_full_search",
        "minoccur")]


def default_params(**kw) -> Params:
    """Defaults of Runopts::validate (src/sortmerna/options.cpp:1684-1738)."""
    p = Params(match=2, mismatch=-3, score_N=-3, gap_open=5, gap_ext=2, num_seeds=2, min_lis=2, edges=4,
               edges_is_percent=0, num_alignments=1, is_best=1, is_forward=1, is_reverse=1, is_full_search=0,
               minoccur=0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


RESULT_DTYPE = np.dtype([("lastIndex", "<u4"), ("lastPart", "<u4"), ("hit_seeds", "<u4"), ("min_index", "<u4"),
                         ("max_index", "<u4"), ("n_align", "<u4"), ("max_SW_count", "<u2"), ("is_done", "u1"),
                         ("is_hit", "u1")])
ALN_DTYPE = np.dtype([("cigar_off", "<u4"), ("cigar_len", "<u4"), ("ref_num", "<u4"), ("ref_begin1", "<i4"),
                      ("ref_end1", "<i4"), ("read_begin1", "<i4"), ("read_end1", "<i4"), ("readlen", "<u4"),
                      ("score1", "<u2"), ("part", "<u2"), ("index_num", "<u2"), ("strand", "u1"), ("pad", "u1")])
COUNTER_NAMES = ("num_aligned", "num_short_last", "sw_calls", "sw_cells", "windows", "trie_nodes",
                 "bucket_entries", "buckets", "pos_entries")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", HERE, "libsmr_oracle.so"])
        L = C.CDLL(LIB_PATH)
        L.ora_index_load.restype = C.c_void_p
        L.ora_index_load.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
        L.ora_index_free.argtypes = [C.c_void_p]
        L.ora_index_num_ids.restype = C.c_uint32
        L.ora_index_num_ids.argtypes = [C.c_void_p]
        L.ora_index_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_seed_window.restype = C.c_int
        L.ora_seed_window.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ora_ssw_align.restype = C.c_int
        L.ora_ssw_align.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_uint16, C.c_void_p, C.c_void_p, C.c_int32]
        L.ora_align.restype = C.c_int
        _lib = L
    return _lib


class OracleIndex:
    def __init__(self, prefix: str, part: int = 0, lnwin: int = 18):
        err = C.create_string_buffer(512)
        self.h = lib().ora_index_load(prefix.encode(), part, lnwin, err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.prefix, self.part, self.lnwin = prefix, part, lnwin

    def stats(self):
        out = np.zeros(8, np.uint64)
        lib().ora_index_stats(self.h, out.ctypes.data)
        return dict(zip(("kmers", "nodes", "buckets", "entries", "ids", "positions", "max_bucket", "max_pos"), map(int, out)))

    def seed_window(self, seq03: np.ndarray, win_pos: int, full_search=False, minoccur=0, cap=4096):
        ids = np.zeros(cap, np.uint32)
        az = C.c_int(0)
        seq03 = np.ascontiguousarray(seq03, np.uint8)
        n = lib().ora_seed_window(self.h, seq03.ctypes.data, win_pos, int(full_search), minoccur, ids.ctypes.data, cap, C.byref(az))
        return ids[:n].copy(), bool(az.value)

    def close(self):
        if self.h:
            lib().ora_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def score_matrix(match=2, mismatch=-3, score_N=-3) -> np.ndarray:
    """Read::initScoringMatrix (read.cpp:274-288), 5x5 row-major."""
    m = np.full((5, 5), score_N, np.int8)
    for a in range(4):
        for b in range(4):
            m[a, b] = match if a == b else mismatch
    return m.reshape(-1).copy()


def ssw_align(read: np.ndarray, ref: np.ndarray, mat: np.ndarray, gap_open=5, gap_ext=2, filters=0):
    read = np.ascontiguousarray(read, np.int8)
    ref = np.ascontiguousarray(ref, np.int8)
    out = np.zeros(6, np.int32)
    cig = np.zeros(len(read) + len(ref) + 8, np.uint32)
    rc = lib().ora_ssw_align(read.ctypes.data, len(read), ref.ctypes.data, len(ref), mat.ctypes.data, gap_open, gap_ext,
                             filters, out.ctypes.data, cig.ctypes.data, len(cig))
    return rc, out, cig[: out[5]].copy()


def align(indexes, index_nums, parts, n_index_files, refs, minimal_scores, skiplengths, params: Params, batch,
          nthreads: int = 1, cigar_cap: int | None = None):
    """ora_align wrapper.  indexes: list[OracleIndex]; refs: list[hostio.References] (one per (index,part))."""
    L = lib()
    nidx = len(indexes)
    n = batch.n
    cat = np.ascontiguousarray(batch.cat, np.uint8)
    off = np.ascontiguousarray(batch.off, np.uint64)
    idx_arr = (C.c_void_p * nidx)(*[ix.h for ix in indexes])
    inum = np.asarray(index_nums, np.uint16)
    ipart = np.asarray(parts, np.uint16)
    refseq = (C.c_void_p * nidx)(*[r.cat.ctypes.data for r in refs])
    refoff = (C.c_void_p * nidx)(*[r.off.ctypes.data for r in refs])
    nref = np.asarray([r.n for r in refs], np.uint32)
    ms = np.asarray(minimal_scores, np.uint32)
    sk = np.asarray(skiplengths, np.uint32).reshape(-1)
    assert sk.size == 3 * nidx
    slots = params.num_alignments if params.num_alignments > 0 else 16   # 0 = all alignments: stride grown on demand
    user_cap = cigar_cap
    L.ora_aln_slots_needed.restype = C.c_uint32
    while True:
        L.ora_set_aln_slots(C.c_uint32(slots))
        res = np.zeros(n, RESULT_DTYPE)
        alns = np.zeros(n * slots, ALN_DTYPE)
        cigar_cap = user_cap if user_cap is not None else 64 * n * slots + 1024
        pool = np.zeros(cigar_cap, np.uint32)
        used = C.c_uint64(0)
        matched = np.zeros(n_index_files, np.uint64)
        counters = np.zeros(len(COUNTER_NAMES), np.uint64)
        rc = L.ora_align(idx_arr, C.c_void_p(inum.ctypes.data), C.c_void_p(ipart.ctypes.data), C.c_uint32(nidx),
                         C.c_uint32(n_index_files), refseq, refoff, C.c_void_p(nref.ctypes.data),
                         C.c_void_p(ms.ctypes.data), C.c_void_p(sk.ctypes.data), C.byref(params),
                         C.c_void_p(cat.ctypes.data), C.c_void_p(off.ctypes.data), C.c_uint32(n),
                         C.c_void_p(res.ctypes.data), C.c_void_p(alns.ctypes.data), C.c_void_p(pool.ctypes.data),
                         C.c_uint64(cigar_cap), C.byref(used), C.c_void_p(matched.ctypes.data),
                         C.c_void_p(counters.ctypes.data), C.c_int(nthreads))
        if rc == 2 and params.num_alignments == 0 and int(L.ora_aln_slots_needed()) > slots:
            slots = int(L.ora_aln_slots_needed())
            continue
        break
    if rc != 0:
        raise RuntimeError(f"ora_align rc={rc}")
    return dict(res=res, alns=alns, cigar=pool[: used.value].copy(), matched=matched,
                counters=dict(zip(COUNTER_NAMES, map(int, counters))), slots=slots)


# ------------------------------------------------------------------------------------------------
# the unmodified reference binary (oracle/_ref/sortmerna_ref)
# ------------------------------------------------------------------------------------------------
def have_reference_binary() -> bool:
    return os.path.exists(REF_BIN)


def run_reference(ref_fastas, reads, workdir, extra=(), threads=1, idx_dir=None, task=4, timeout=3600):
    """Run the reference CLI; returns dict(log=<aligned.log text>, stdout=..., out_dir=..., idx_dir=...)."""
    cmd = [REF_BIN]
    for r in ref_fastas:
        cmd += ["-ref", r]
    for r in ([reads] if isinstance(reads, str) else reads):
        cmd += ["-reads", r]
    cmd += ["-workdir", workdir, "-threads", str(threads), "-task", str(task)]
    if idx_dir:
        cmd += ["-idx-dir", idx_dir]
    cmd += list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference failed ({p.returncode}): {' '.join(cmd)}\n{p.stdout[-3000:]}")
    out_dir = os.path.join(workdir, "out")
    log = ""
    lp = os.path.join(out_dir, "aligned.log")
    if os.path.exists(lp):
        log = open(lp).read()
    return dict(log=log, stdout=p.stdout, out_dir=out_dir, idx_dir=idx_dir or os.path.join(workdir, "idx"), cmd=cmd)


def parse_log(log: str) -> dict:
    """Numbers the integration tests of the reference read back from aligned.log (scripts/run.py:245-273)."""
    d = dict(lambda_=[], K=[], minimal_score=[])
    for m in re.finditer(r"Gumbel lambda = ([0-9.eE+-]+)", log):
        d["lambda_"].append(float(m.group(1)))
    for m in re.finditer(r"Gumbel K = ([0-9.eE+-]+)", log):
        d["K"].append(float(m.group(1)))
    for m in re.finditer(r"Minimal SW score based on E-value = (\d+)", log):
        d["minimal_score"].append(int(m.group(1)))
    m = re.search(r"Total reads = (\d+)", log)
    d["total_reads"] = int(m.group(1)) if m else None
    m = re.search(r"Total reads passing E-value threshold = (\d+)", log)
    d["passing"] = int(m.group(1)) if m else None
    m = re.search(r"Total reads failing E-value threshold = (\d+)", log)
    d["failing"] = int(m.group(1)) if m else None
    d["coverage"] = [float(x) for x in re.findall(r"\t\t([0-9.]+)\s*$", log, re.M)]
    return d


def read_sam_rows(path: str) -> list:
    return [ln.rstrip("\n") for ln in open(path) if not ln.startswith("@")]
