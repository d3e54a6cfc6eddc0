"""Differential fuzzing of the oracle against the UNMODIFIED reference binary: random database slices, damaged reads and
random option sets (scoring, -num_alignments, -no-best, -F/-R, -full_search, -num_seeds, -min_lis, -edges): identical SAM rows
and pass/fail totals.  (The GPU twin of this test, oracle vs kernels on the same generator, is tests/test_gpu_fuzz.py.)"""
import os
import shutil
import tempfile

import pytest

from conftest import ROOT
from fuzz_common import make_case
from helpers import params_kwargs_from_args, strip_seq
from sortmerna_b200 import api, hostio

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sortmerna_ref")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/sortmerna_ref not built")
@pytest.mark.parametrize("seed", range(int(os.environ.get("SMR_FUZZ_SEEDS", "12"))))
def test_oracle_equals_reference_on_random_cases(seed):
    from oracle import ora
    d = tempfile.mkdtemp(prefix="smr_fz_")
    try:
        fastas, reads_p, args = make_case(seed, d)
        try:
            r = ora.run_reference(fastas, reads_p, os.path.join(d, "w"), extra=["-sam", "-fastx"] + args, threads=1)
        except RuntimeError as e:
            if "Sls::error" in str(e) or "ALP" in str(e):
                pytest.skip("scoring set outside the reference's Gumbel tables")
            if "[validate:" in str(e):
                pytest.skip("option combination refused by the reference's own validation")
            raise
        log = ora.parse_log(r["log"])
        sam = strip_seq(ora.read_sam_rows(os.path.join(r["out_dir"], "aligned.sam")))
        batch = hostio.load_reads(reads_p)
        refs = [hostio.load_references(f) for f in fastas]
        kw = params_kwargs_from_args(args)
        oix, prefixes = [], []
        for k, f in enumerate(fastas):
            p = os.path.join(d, f"idx{k}")
            api.build_index(f, p)                       # our builder (equal to the reference's: tests/test_index_builder.py)
            oix.append(ora.OracleIndex(p, 0, 18))
        out = ora.align(oix, [0, 1], [0, 0], 2, refs, log["minimal_score"], [18, 9, 3, 18, 9, 3], ora.default_params(**kw), batch, nthreads=2)
        rows = strip_seq(hostio.format_sam_rows(batch, refs, out["res"], out["alns"], out["cigar"], out["slots"]))
        assert sorted(rows) == sorted(sam), args
        assert int(out["res"]["is_hit"].sum()) == log["passing"], args
    finally:
        shutil.rmtree(d, ignore_errors=True)


INDEX_OPTS = [(["-max_pos", "2"], dict(max_pos=2)), (["-max_pos", "7"], dict(max_pos=7)), (["-interval", "2"], dict(interval=2)),
              (["-m", "0.4"], dict(max_mb=0.4)), (["-max_pos", "3", "-m", "0.25"], dict(max_pos=3, max_mb=0.25))]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/sortmerna_ref not built")
@pytest.mark.parametrize("seed", range(int(os.environ.get("SMR_FUZZ_INDEX_SEEDS", "6"))))
def test_oracle_equals_reference_with_index_options(seed):
    """The same differential test with index-build options in play: truncated position lists (-max_pos), sparse (L+1)-mers
    (-interval) and indexes split into several parts (-m): the oracle runs on the files smr_build_index writes, the reference on
    the ones its own builder writes."""
    from oracle import ora
    d = tempfile.mkdtemp(prefix="smr_fzi_")
    try:
        fastas, reads_p, args = make_case(5000 + seed, d)
        ref_extra, kw_build = INDEX_OPTS[seed % len(INDEX_OPTS)]
        try:
            r = ora.run_reference(fastas, reads_p, os.path.join(d, "w"), extra=["-sam", "-fastx"] + args + ref_extra, threads=1)
        except RuntimeError as e:
            if "[validate:" in str(e) or "Sls::error" in str(e):
                pytest.skip("option combination refused by the reference")
            raise
        log = ora.parse_log(r["log"])
        sam = strip_seq(ora.read_sam_rows(os.path.join(r["out_dir"], "aligned.sam")))
        batch = hostio.load_reads(reads_p)
        kw = params_kwargs_from_args(args)
        oix, inum, parts, refs, ms, by_index = [], [], [], [], [], []
        for k, f in enumerate(fastas):
            p = os.path.join(d, f"idx{k}")
            api.build_index(f, p, **kw_build)
            st = hostio.parse_stats(p)
            pr = hostio.split_by_parts(hostio.load_references(f), st)
            by_index.append(pr)
            for part in range(st.num_parts):
                oix.append(ora.OracleIndex(p, part, 18)); inum.append(k); parts.append(part); refs.append(pr[part]); ms.append(log["minimal_score"][k])
        out = ora.align(oix, inum, parts, 2, refs, ms, [18, 9, 3] * len(oix), ora.default_params(**kw), batch, nthreads=2)
        rows = strip_seq(hostio.format_sam_rows(batch, by_index, out["res"], out["alns"], out["cigar"], out["slots"]))
        assert sorted(rows) == sorted(sam), (args, ref_extra)
        assert int(out["res"]["is_hit"].sum()) == log["passing"]
    finally:
        shutil.rmtree(d, ignore_errors=True)
