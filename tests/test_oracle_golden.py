"""The oracle (oracle/smr_oracle.cpp) against what the unmodified reference binary printed
(tests/golden/case_*/expected.json, made by tests/golden/make_golden.py): every SAM row
(read, strand flag, reference, position, CIGAR with soft clips, AS:i score, NM:i) must be identical,
and so must the pass/fail totals and the per-database coverage of aligned.log."""
import os

import numpy as np
import pytest

from conftest import case_names, load_case, load_denovo
from helpers import assert_blast_rows_equal, blast_rows, params_kwargs_from_args, strip_seq
from oracle import ora
from sortmerna_b200 import hostio


@pytest.fixture(scope="module")
def oracle_indexes(golden):
    return [ora.OracleIndex(p, 0, s.lnwin) for p, s in zip(golden["prefixes"], golden["stats"])]


def run_oracle(golden, oracle_indexes, case, nthreads=2):
    exp = load_case(case)
    prm = ora.default_params(**params_kwargs_from_args(exp["args"]))
    out = ora.align(oracle_indexes, [0, 1], [0, 0], 2, golden["refs"], exp["log"]["minimal_score"], [18, 9, 3, 18, 9, 3], prm,
                    golden["batch"], nthreads=nthreads)
    return exp, prm, out


@pytest.mark.parametrize("case", case_names())
def test_oracle_matches_reference_sam(golden, oracle_indexes, case):
    exp, prm, out = run_oracle(golden, oracle_indexes, case)
    rows = hostio.format_sam_rows(golden["batch"], golden["refs"], out["res"], out["alns"], out["cigar"], out["slots"])
    if case != "default":
        rows = strip_seq(rows)
    # with several references the reference writes rows index-major; compare as multisets
    assert sorted(rows) == sorted(exp["sam"])
    assert int(out["res"]["is_hit"].sum()) == exp["log"]["passing"]
    assert golden["batch"].n - int(out["res"]["is_hit"].sum()) == exp["log"]["failing"]
    # aligned.log "Coverage by database" (summary.cpp:160-170), %.2f of matched/total
    cov = [round(100.0 * int(m) / golden["batch"].n, 2) for m in out["matched"]]
    assert cov == pytest.approx(exp["log"]["coverage"], abs=0.011)


@pytest.mark.parametrize("case", case_names())
def test_report_arithmetic_matches_reference_blast(golden, oracle_indexes, case):
    """calc_miss_gap_match / %id / %cov / bitscore / E-value / clipped CIGAR restated in hostio reproduce the reference's BLAST rows."""
    exp, prm, out = run_oracle(golden, oracle_indexes, case)
    st = hostio.host_aln_stats(golden["batch"], golden["refs"], out["res"], out["alns"], out["cigar"], out["slots"])
    assert_blast_rows_equal(blast_rows(golden, exp, out, st), exp["blast"])


@pytest.mark.parametrize("case", ["default", "best3", "rev_only", "loose"])
def test_denovo_classification_matches_reference(golden, oracle_indexes, case):
    """denovo_stats counters and the aligned_denovo read set (-otu_map -de_novo_otu -id X -coverage Y) from the host restatement."""
    dn = load_denovo()[case]
    prm = ora.default_params(**params_kwargs_from_args(dn["args"]))
    b = golden["batch"]
    out = ora.align(oracle_indexes, [0, 1], [0, 0], 2, golden["refs"], dn["minimal_score"], [18, 9, 3, 18, 9, 3], prm, b, nthreads=2)
    st = hostio.host_aln_stats(b, golden["refs"], out["res"], out["alns"], out["cigar"], out["slots"])
    cls = hostio.denovo_classes(out["res"], out["alns"], out["slots"], st, dn["min_id"], dn["min_cov"])
    assert cls.sum(axis=0).tolist() == dn["counts"]
    assert int(cls[:, 3].sum()) == dn["total_denovo"]
    ids = sorted(hostio.seq_id(b.headers[r]) for r in np.nonzero(hostio.is_denovo_read(cls))[0])
    assert ids == dn["denovo_reads"]


def test_oracle_multipart_index_matches_reference(golden, golden_parts):
    """Index split into 3 parts per database (-m 0.5): the per-part loop of align() (processor.cpp:196-262) -- part-relative
    ref_num, Read::best re-initialised per part, stored state carried across parts -- against the reference's own run."""
    exp = load_case("parts")
    assert [g["stats"].num_parts for g in golden_parts] == exp["num_parts"] == [3, 3]
    oix, inum, parts, refs, ms = [], [], [], [], []
    for k, g in enumerate(golden_parts):
        for p in range(g["stats"].num_parts):
            oix.append(ora.OracleIndex(g["prefix"], p, g["stats"].lnwin)); inum.append(k); parts.append(p)
            refs.append(g["part_refs"][p]); ms.append(exp["log"]["minimal_score"][k])
    out = ora.align(oix, inum, parts, 2, refs, ms, [18, 9, 3] * len(oix), ora.default_params(), golden["batch"], nthreads=2)
    by_index = [g["part_refs"] for g in golden_parts]
    rows = strip_seq(hostio.format_sam_rows(golden["batch"], by_index, out["res"], out["alns"], out["cigar"], out["slots"]))
    assert sorted(rows) == sorted(exp["sam"])
    assert int(out["res"]["is_hit"].sum()) == exp["log"]["passing"]
    st = hostio.host_aln_stats(golden["batch"], by_index, out["res"], out["alns"], out["cigar"], out["slots"])
    b = golden["batch"]
    tot = int(np.diff(b.off.astype(np.int64)).sum())
    gum = list(zip(exp["log"]["lambda_"], exp["log"]["K"]))
    evp = [hostio.evalue_params(g["stats"], k, tot, b.n) for g, (_, k) in zip(golden_parts, gum)]
    assert_blast_rows_equal(hostio.format_blast_rows(b, by_index, out["res"], out["alns"], out["cigar"], out["slots"], st, gum, evp), exp["blast"])


def test_oracle_thread_invariance(golden, oracle_indexes):
    _, _, a = run_oracle(golden, oracle_indexes, "default", nthreads=1)
    _, _, b = run_oracle(golden, oracle_indexes, "default", nthreads=4)
    assert np.array_equal(a["res"], b["res"]) and np.array_equal(a["alns"], b["alns"]) and np.array_equal(a["cigar"], b["cigar"])


def test_minimal_score_formula(golden):
    """refstats.cpp:236-265 restated in hostio.minimal_score reproduces the reference's log."""
    exp = load_case("default")
    b = golden["batch"]
    lens = np.diff(b.off.astype(np.int64))
    for k, st in enumerate(golden["stats"]):
        ms = hostio.minimal_score(st, exp["log"]["lambda_"][k], exp["log"]["K"][k], int(lens.sum()), b.n)
        assert ms == exp["log"]["minimal_score"][k]


def test_empty_and_short_batches(golden, oracle_indexes):
    prm = ora.default_params()
    batch = hostio.pack_reads(["@e", "@s", "@x"], [b"", b"ACGTACGTAC", b"ACGTNNNN"])
    out = ora.align(oracle_indexes, [0, 1], [0, 0], 2, golden["refs"], [37, 36], [18, 9, 3, 18, 9, 3], prm, batch)
    assert int(out["res"]["is_hit"].sum()) == 0 and out["counters"]["num_short_last"] == 3


def test_oracle_long_read_t0(golden_t0):
    """BASELINE config 1 (t0/t2 of the reference's suite): 1.5 kb read, int16 'word' SW path, 30-op CIGAR."""
    g = golden_t0
    ix = ora.OracleIndex(g["prefix"], 0, g["stats"].lnwin)
    out = ora.align([ix], [0], [0], 1, [g["refs"]], g["exp"]["log"]["minimal_score"], [18, 9, 3], ora.default_params(), g["batch"])
    rows = hostio.format_sam_rows(g["batch"], [g["refs"]], out["res"], out["alns"], out["cigar"], out["slots"])
    assert rows == g["exp"]["sam"]
    assert int(out["alns"]["score1"][0]) > 255   # the byte kernel overflowed in the reference: word path
