"""Host-side I/O helpers around the alignment hot path (numpy only; no GPU, no oracle).

They restate the small pieces of the reference that sit immediately on either side of the
drop-in boundary so that tests and bench.py can feed the C-ABI the same bytes the reference's
`align()` sees:

  * `encode_nt`      -- `nt_table` (include/common.hpp:68-77): ACGTU -> 0..3, anything else -> 4
  * `load_references`-- `References::load` + `convert_fix` (src/sortmerna/references.cpp:55-164)
  * `read_fastx`     -- single-line-record FASTA/FASTQ reader (the reference's Readfeed is out of scope)
  * `parse_stats`    -- the `.stats` index file (`Refstats::load`, src/sortmerna/refstats.cpp:103-190)
  * `minimal_score`  -- the E-value -> minimal SW score formula (refstats.cpp:236-265) given lambda, K
  * `format_sam_rows`-- the SAM row layout of `ReportSam::append` (src/sortmerna/report_sam.cpp:64-152)
"""
from __future__ import annotations

import glob
import gzip
import math
import os
import struct
from dataclasses import dataclass, field

import numpy as np

_NT = np.full(256, 4, dtype=np.uint8)
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _NT[ord(_c)] = _v
    _NT[ord(_c.lower())] = _v
NT_MAP = "ACGTN"


def encode_nt(seq: bytes | str) -> np.ndarray:
    """ASCII -> 0..4 (4 = ambiguous), include/common.hpp:68-77."""
    if isinstance(seq, str):
        seq = seq.encode()
    return _NT[np.frombuffer(seq, dtype=np.uint8)]


def _open(path: str):
    return gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")


def read_fastx(path: str, limit: int | None = None):
    """Return (headers, seqs, quals) of a FASTA/FASTQ file (multi-line FASTA is concatenated)."""
    headers, seqs, quals = [], [], []
    with _open(path) as fh:
        data = fh.read()
    lines = data.split(b"\n")
    i, n = 0, len(lines)
    while i < n:
        ln = lines[i].rstrip(b"\r")
        if not ln:
            i += 1
            continue
        if ln[:1] == b"@":
            headers.append(ln.decode())
            seqs.append(lines[i + 1].rstrip(b"\r"))
            quals.append(lines[i + 3].rstrip(b"\r"))
            i += 4
        elif ln[:1] == b">":
            headers.append(ln.decode())
            i += 1
            parts = []
            while i < n and lines[i][:1] != b">":
                if lines[i].strip():
                    parts.append(lines[i].strip())
                i += 1
            seqs.append(b"".join(parts))
            quals.append(b"")
        else:
            raise ValueError(f"{path}: unexpected line {i}: {ln[:40]!r}")
        if limit is not None and len(seqs) >= limit:
            break
    return headers, seqs, quals


def seq_id(header: str) -> str:
    """Read::getSeqId (read.cpp:365-371): header up to the first space, leading '>'/'@' removed."""
    h = header.split(" ")[0]
    return h.lstrip(">@")


@dataclass
class ReadBatch:
    headers: list
    seqs: list            # raw bytes
    quals: list
    cat: np.ndarray       # uint8, 0..4
    off: np.ndarray       # uint64, nreads+1

    @property
    def n(self):
        return len(self.seqs)


def pack_reads(headers, seqs, quals=None) -> ReadBatch:
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    cat = encode_nt(b"".join(seqs)) if seqs else np.zeros(0, np.uint8)
    return ReadBatch(list(headers), list(seqs), list(quals) if quals else [b""] * len(seqs), cat, off)


def load_reads(path: str, limit: int | None = None) -> ReadBatch:
    h, s, q = read_fastx(path, limit)
    return pack_reads(h, s, q)


@dataclass
class References:
    path: str
    ids: list             # BaseRecord::getId -- header up to first space without '>'
    cat: np.ndarray       # uint8 0..4, all sequences concatenated
    off: np.ndarray       # uint64, nref+1

    @property
    def n(self):
        return len(self.ids)


def load_references(path: str) -> References:
    """References::load for a single-part index (references.cpp:55-154)."""
    h, s, _ = read_fastx(path)
    lens = np.fromiter((len(x) for x in s), dtype=np.uint64, count=len(s))
    off = np.zeros(len(s) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return References(path, [seq_id(x) for x in h], encode_nt(b"".join(s)), off)


def split_by_parts(refs: "References", stats: "IndexStats") -> list:
    """The references of each index part (References::load reads [start_part, start_part + seq_part_size) of the FASTA,
    references.cpp:55-154): `ref_num` of an alignment is relative to its part."""
    out, first = [], 0
    for (_, _, nseq) in stats.parts:
        off = refs.off[first:first + nseq + 1]
        out.append(References(refs.path, refs.ids[first:first + nseq], refs.cat[int(off[0]):int(off[-1])], (off - off[0]).astype(np.uint64)))
        first += nseq
    return out


def _refs_of(refs_by_index, al):
    """refs_by_index[index_num] is a References (single-part index) or the list split_by_parts returns"""
    refs = refs_by_index[int(al["index_num"])]
    return refs[int(al["part"])] if isinstance(refs, (list, tuple)) else refs


@dataclass
class IndexStats:
    """Contents of <prefix>.stats (refstats.cpp:129-190; writer indexdb.cpp:2020-2080)."""
    prefix: str
    fasta_size: int = 0
    fasta_name: str = ""
    background_freq: tuple = (0.25, 0.25, 0.25, 0.25)
    full_ref: int = 0
    lnwin: int = 18
    numseq: int = 0
    num_parts: int = 1
    parts: list = field(default_factory=list)  # (start_part, seq_part_size, numseq_part)


def parse_stats(prefix: str) -> IndexStats:
    with open(prefix + ".stats", "rb") as fh:
        b = fh.read()
    o = 0
    (fsize,) = struct.unpack_from("<Q", b, o); o += 8
    (nlen,) = struct.unpack_from("<I", b, o); o += 4
    name = b[o:o + nlen].split(b"\0")[0].decode(); o += nlen
    freq = struct.unpack_from("<4d", b, o); o += 32
    (full_ref,) = struct.unpack_from("<Q", b, o); o += 8
    (lnwin,) = struct.unpack_from("<I", b, o); o += 4
    (numseq,) = struct.unpack_from("<Q", b, o); o += 8
    (nparts,) = struct.unpack_from("<H", b, o); o += 2
    parts = []
    for _ in range(nparts):
        sp, sz, ns = struct.unpack_from("<QQI", b, o); o += 24  # struct index_parts_stats, 8+8+4(+4 pad)
        parts.append((sp, sz, ns))
    return IndexStats(prefix, fsize, name, freq, full_ref, lnwin, numseq, nparts, parts)


def find_index_prefixes(idx_dir: str) -> dict:
    """Map reference FASTA basename -> index prefix for every *.stats in idx_dir."""
    out = {}
    for st in glob.glob(os.path.join(idx_dir, "*.stats")):
        s = parse_stats(st[: -len(".stats")])
        out[os.path.basename(s.fasta_name)] = s.prefix
    return out


def minimal_score(stats: IndexStats, lam: float, K: float, all_reads_len: int, all_reads_count: int,
                  evalue: float = 1.0) -> int:
    """refstats.cpp:236-265 (is_score_split = false)."""
    f = stats.background_freq
    entropy = -sum(p * math.log2(p) for p in f)
    full_ref = stats.full_ref
    full_read = all_reads_len
    expect_L = int(math.log(K * full_ref * full_read) / entropy)
    if full_ref > expect_L * stats.numseq:
        full_ref -= expect_L * stats.numseq
    full_read -= expect_L * all_reads_count
    return int(math.log(evalue / (K * full_ref * full_read)) / -lam) & 0xFFFFFFFF


def evalue_params(stats: IndexStats, K: float, all_reads_len: int, all_reads_count: int) -> tuple:
    """the length-corrected (full_ref, full_read) of refstats.cpp:236-257 that the E-value of a BLAST row uses"""
    entropy = -sum(p * math.log2(p) for p in stats.background_freq)
    full_ref, full_read = stats.full_ref, all_reads_len
    expect_L = int(math.log(K * full_ref * full_read) / entropy)
    if full_ref > expect_L * stats.numseq:
        full_ref -= expect_L * stats.numseq
    full_read -= expect_L * all_reads_count
    return full_ref, full_read


def _g3(x: float) -> str:
    """what `ss.precision(3); ss << x` prints (C++ general format, 3 significant digits)"""
    return f"{x:.3g}"


def format_blast_rows(batch: "ReadBatch", refs_by_index: list, results, alns, cigar_pool, slots: int, stats, gumbel: list,
                      ev_params: list) -> list:
    """Tabular BLAST rows with the optional columns 'cigar qcov qstrand' as ReportBlast::append prints them
    (src/sortmerna/report_blast.cpp:99-365).  stats[i] = (n_miss, n_gap, n_match) of alignment i -- from the GPU
    (smr_aln_stats) or from calc_miss_gap_match; gumbel[index] = (lambda, K); ev_params[index] = (full_ref, full_read)."""
    import numpy as _np
    rows = []
    for r in range(batch.n):
        na = int(results["n_align"][r])
        name = seq_id(batch.headers[r])
        rlen = len(batch.seqs[r])
        for a in range(na):
            al = alns[r * slots + a]
            st = stats[r * slots + a]
            idx = int(al["index_num"])
            lam, K = gumbel[idx]
            full_ref, full_read = ev_params[idx]
            score = int(al["score1"])
            bitscore = int(_np.float32(_np.float32(lam * score - math.log(K)) / _np.float32(math.log(2))))   # report_blast.cpp:117-119
            evalue = K * full_ref * full_read * math.exp(-lam * score)                                      # :121-126
            miss, gap, match = int(st["n_miss"]), int(st["n_gap"]), int(st["n_match"])
            pid = match / (miss + gap + match)
            cov = abs(int(al["read_end1"]) - int(al["read_begin1"]) + 1) / int(al["readlen"])
            cig = cigar_pool[int(al["cigar_off"]):int(al["cigar_off"]) + int(al["cigar_len"])]
            cs = (f"{int(al['read_begin1'])}S" if int(al["read_begin1"]) else "") + cigar_string(cig)
            end_mask = rlen - int(al["read_end1"]) - 1
            if end_mask > 0:
                cs += f"{end_mask}S"
            refs = _refs_of(refs_by_index, al)
            rows.append("\t".join([name, refs.ids[int(al["ref_num"])], _g3(pid * 100), str(int(al["read_end1"]) - int(al["read_begin1"]) + 1),
                                   str(miss), str(gap), str(int(al["read_begin1"]) + 1), str(int(al["read_end1"]) + 1),
                                   str(int(al["ref_begin1"]) + 1), str(int(al["ref_end1"]) + 1), _g3(evalue), str(bitscore), cs,
                                   _g3(cov * 100), "+" if bool(al["strand"]) else "-"]))
    return rows


def host_aln_stats(batch: "ReadBatch", refs_by_index: list, results, alns, cigar_pool, slots: int):
    """calc_miss_gap_match on the host (numpy) for every stored alignment: the CPU twin of smr_aln_stats, used by the tests."""
    out = np.zeros(batch.n * slots, dtype=[("n_miss", "<u4"), ("n_gap", "<u4"), ("n_match", "<u4"), ("n_match_denovo", "<u4")])
    for r in range(batch.n):
        enc = batch.cat[int(batch.off[r]):int(batch.off[r + 1])]
        for a in range(int(results["n_align"][r])):
            al = alns[r * slots + a]
            refs = _refs_of(refs_by_index, al)
            e04 = enc if bool(al["strand"]) else np.where(enc < 4, 3 - enc, 4)[::-1]
            rseq = refs.cat[int(refs.off[int(al["ref_num"])]):int(refs.off[int(al["ref_num"]) + 1])]
            cig = cigar_pool[int(al["cigar_off"]):int(al["cigar_off"]) + int(al["cigar_len"])]
            m = calc_miss_gap_match(rseq, e04, al, cig)
            # denovo_stats_run (processor.cpp:329-357) walks the same CIGAR over the read WITHOUT reverse-complementing it
            md = m if bool(al["strand"]) else calc_miss_gap_match(rseq, enc, al, cig)
            out[r * slots + a] = (m[0], m[1], m[2], md[2])
    return out


def denovo_classes(results, alns, slots: int, stats, min_id: float, min_cov: float):
    """denovo_stats_run (processor.cpp:329-357) from smr_aln_stats: per read the four counters
    (c_yid_ycov, n_yid_ncov, n_nid_ycov, n_denovo) over its stored alignments; their column sums are Readstats'
    n_yid_ycov / n_yid_ncov / n_nid_ycov / num_denovo.  Returns an (nreads, 4) uint32 array."""
    n = results.shape[0]
    out = np.zeros((n, 4), np.uint32)
    for r in range(n):
        for a in range(int(results["n_align"][r])):
            al, st = alns[r * slots + a], stats[r * slots + a]
            tot = int(st["n_miss"]) + int(st["n_gap"]) + int(st["n_match"])
            idv = int(st["n_match_denovo"]) / tot
            cov = abs(int(al["read_end1"]) - int(al["read_begin1"]) + 1) / int(al["readlen"])
            is_id = math.floor(idv * 1000.0 + 0.5) / 1000.0 >= min_id      # :336-339 round to 3 decimals
            is_cov = math.floor(cov * 1000.0 + 0.5) / 1000.0 >= min_cov
            out[r, 0 if (is_id and is_cov) else 1 if is_id else 2 if is_cov else 3] += 1
    return out


def is_denovo_read(classes: np.ndarray) -> np.ndarray:
    """output.cpp:130-141 (single-end): the read goes to aligned_denovo.* when n_denovo > 0 and the other three are 0"""
    return (classes[:, 3] > 0) & (classes[:, 0] == 0) & (classes[:, 1] == 0) & (classes[:, 2] == 0)


_OPS = "MID"


def cigar_string(cig) -> str:
    return "".join(f"{int(c) >> 4}{_OPS[int(c) & 0xF] if (int(c) & 0xF) < 3 else 'D'}" for c in cig)


def revcomp_str(s: str) -> str:
    return s.translate(str.maketrans("ACGTN", "TGCAN"))[::-1]


def calc_miss_gap_match(ref: np.ndarray, read04: np.ndarray, aln, cigar) -> tuple:
    """Read::calc_miss_gap_match (read.cpp:547-589): (mismatches, gaps, matches)."""
    qb, pb = int(aln["ref_begin1"]), int(aln["read_begin1"])
    miss = gap = match = 0
    for c in cigar:
        op, ln = int(c) & 0xF, int(c) >> 4
        if op == 0:
            a = ref[qb:qb + ln]
            b = read04[pb:pb + ln]
            eq = int(np.count_nonzero(a == b))
            match += eq
            miss += ln - eq
            qb += ln
            pb += ln
        elif op == 1:
            pb += ln
            gap += ln
        else:
            qb += ln
            gap += ln
    return miss, gap, match


def format_sam_rows(batch: ReadBatch, refs_by_index: list, results, alns, cigar_pool, slots: int,
                    with_seq: bool = True) -> list:
    """SAM alignment rows as `ReportSam::append` prints them (report_sam.cpp:64-152), one list entry
    per stored alignment, in read order.  `results`/`alns` are the structured numpy arrays returned
    by the C-ABI (or the oracle); refs_by_index[index_num] is a `References`."""
    rows = []
    for r in range(batch.n):
        na = int(results["n_align"][r])
        if na == 0:
            continue
        seq = batch.seqs[r].decode().upper()
        enc = batch.cat[int(batch.off[r]):int(batch.off[r + 1])]
        name = seq_id(batch.headers[r])
        for a in range(na):
            al = alns[r * slots + a]
            refs = _refs_of(refs_by_index, al)
            cig = cigar_pool[int(al["cigar_off"]):int(al["cigar_off"]) + int(al["cigar_len"])]
            strand = bool(al["strand"])
            cs = ""
            if int(al["read_begin1"]) != 0:
                cs += f"{int(al['read_begin1'])}S"
            cs += cigar_string(cig)
            end_mask = len(seq) - int(al["read_end1"]) - 1
            if end_mask > 0:
                cs += f"{end_mask}S"
            # the read as aligned (04 alphabet, reverse-complemented for the minus strand)
            s04 = "".join(NT_MAP[v] for v in enc)
            e04 = enc
            if not strand:
                s04 = revcomp_str(s04)
                e04 = np.where(enc < 4, 3 - enc, 4)[::-1]
            rseq = refs.cat[int(refs.off[int(al["ref_num"])]):int(refs.off[int(al["ref_num"]) + 1])]
            miss, gap, _ = calc_miss_gap_match(rseq, e04, al, cig)
            q = batch.quals[r].decode() if batch.quals[r] else "*"
            if batch.quals[r] and not strand:
                q = q[::-1]
            row = [name, "16" if not strand else "0", refs.ids[int(al["ref_num"])], str(int(al["ref_begin1"]) + 1),
                   "255", cs, "*", "0", "0", s04 if with_seq else "*", q if with_seq else "*",
                   f"AS:i:{int(al['score1'])}", f"NM:i:{miss + gap}"]
            rows.append("\t".join(row))
    return rows
