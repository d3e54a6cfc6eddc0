"""The drop-in, end to end on the GPU: the unmodified reference host program linked with integration/align_gpu.cpp and the
PRODUCT library (oracle/_ref/sortmerna_gpu) writes the files the reference binary writes for the same command line."""
import os
import shutil
import tempfile

import pytest

from conftest import GOLDEN
from integration_common import REF_DIR, assert_same_outputs, run_host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("extra", [[], ["-num_alignments", "3"]], ids=["default", "best3"])
def test_reference_host_with_gpu_library(extra):
    for b in ("sortmerna_ref", "sortmerna_gpu"):
        if not os.path.exists(os.path.join(REF_DIR, b)):
            pytest.skip(f"oracle/_ref/{b} not built (oracle/Makefile.ref)")
    d = tempfile.mkdtemp(prefix="smr_integ_gpu_")
    try:
        reads = [os.path.join(GOLDEN, "reads_mix.fq")]
        rep = ["-sam", "-blast", "1 cigar qcov qstrand", "-fastx", "-other"]
        ref, _ = run_host("sortmerna_ref", os.path.join(d, "ref"), reads, rep + extra)
        got, log = run_host("sortmerna_gpu", os.path.join(d, "got"), reads, rep + extra)
        assert "Starting alignment (libsmr_b200)" in log
        assert_same_outputs(got, ref)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_reference_t9_golden_sam_rows_on_gpu():
    """scripts/test.jinja t9 (exact SAM rows, forward + reverse-complement reference) through the drop-in host program"""
    from test_integration_binding import T9_ROWS, run_t9
    if not os.path.exists(os.path.join(REF_DIR, "sortmerna_gpu")):
        pytest.skip("oracle/_ref/sortmerna_gpu not built")
    assert run_t9("sortmerna_gpu") == T9_ROWS


def test_reference_t5_mate_pairs_known_answer():
    """scripts/test.jinja t5: the two set4 mate files (2 x 5000 reads, paired feed, 5 threads) vs silva-bac-16s-id85 built with
    -max_pos 250: 10000 reads -> 6000 aligned / 4000 not; the drop-in host program on the GPU against the reference binary."""
    import subprocess
    from conftest import ROOT
    from tools import stage_data
    cache = os.path.join(ROOT, "data_cache")
    fasta = os.path.join(cache, "sets", "silva-bac-16s-database-id85.fasta")
    reads = [os.path.join(cache, "sets", f"set4_mate_pairs_metatranscriptomics_{k}.fastq") for k in (1, 2)]
    for p in [fasta] + reads + [os.path.join(REF_DIR, "sortmerna_gpu"), os.path.join(REF_DIR, "sortmerna_ref")]:
        if not os.path.exists(p):
            pytest.skip(f"{p} missing")
    idx, _ = stage_data.ensure_indexes([fasta], os.path.join(cache, "idx_set2_ref"), extra=("-max_pos", "250"), builder="reference")
    d = tempfile.mkdtemp(prefix="smr_t5_")
    try:
        outs = {}
        for b in ("sortmerna_ref", "sortmerna_gpu"):
            wd = os.path.join(d, b)
            cmd = [os.path.join(REF_DIR, b), "-ref", fasta, "-reads", reads[0], "-reads", reads[1], "-max_pos", "250", "-fastx", "-other", "-threads", "5",
                   "-workdir", wd, "-idx-dir", idx, "-task", "4"]
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
            assert p.returncode == 0, p.stdout[-2000:]
            log = open(os.path.join(wd, "out", "aligned.log")).read()
            outs[b] = (log, {fn: open(os.path.join(wd, "out", fn)).read() for fn in os.listdir(os.path.join(wd, "out")) if fn.endswith(".fq")})
        for b in outs:
            assert "passing E-value threshold = 6000" in outs[b][0] and "failing E-value threshold = 4000" in outs[b][0], b   # scripts/test.jinja:318-322
        assert outs["sortmerna_gpu"][1] == outs["sortmerna_ref"][1]
    finally:
        shutil.rmtree(d, ignore_errors=True)
