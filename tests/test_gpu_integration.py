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


def _run_binary(binary, fastas, reads, idx, wd, extra, threads=8, env=None):
    import subprocess
    cmd = [os.path.join(REF_DIR, binary)] + sum((["-ref", f] for f in fastas), []) + sum((["-reads", r] for r in reads), []) + \
          ["-workdir", wd, "-idx-dir", idx, "-threads", str(threads), "-task", "4"] + list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1800, env=env)
    assert p.returncode == 0, p.stdout[-3000:]
    out = os.path.join(wd, "out")
    files = {}
    import gzip
    for fn in sorted(os.listdir(out)):
        path = os.path.join(out, fn)
        text = gzip.open(path, "rt", errors="replace").read() if fn.endswith(".gz") else open(path, errors="replace").read()   # gz input -> gz reports
        lines = text.split("\n")
        fn = fn[:-3] if fn.endswith(".gz") else fn
        if fn.endswith(".sam"):
            lines = sorted(ln for ln in lines if ln and not ln.startswith("@"))   # several references: row order depends on the slot count
        elif fn.endswith(".log"):
            lines = [ln for ln in lines if "E-value threshold" in ln or "Total reads =" in ln]
        files[fn] = lines
    return files, p.stdout


def _config4_inputs():
    from conftest import ROOT
    from tools import stage_data
    cache = os.path.join(ROOT, "data_cache")
    reads = [os.path.join(cache, "sets", f"set4_mate_pairs_metatranscriptomics_{k}.fastq.gz") for k in (1, 2)]
    fastas = [stage_data.db_path(n) for n in stage_data.DBS]
    for p in reads + fastas + [os.path.join(REF_DIR, "sortmerna_gpu"), os.path.join(REF_DIR, "sortmerna_ref")]:
        if not os.path.exists(p):
            pytest.skip(f"{p} missing")
    idx, _ = stage_data.ensure_indexes(fastas, os.path.join(cache, "idx_ref"), builder="reference")
    return fastas, reads, idx


def test_baseline_config4_through_the_binary():
    """BASELINE config 4 as written: set4 paired FASTQ.gz vs the 8 rRNA databases, -sam, through the CLI of the drop-in host program
    (gz inflate + paired feed + 8 resident indexes + report stage) against the unmodified reference binary: 5944 / 4056
    (scripts/test.jinja t17) and identical SAM rows / aligned / other reads."""
    fastas, reads, idx = _config4_inputs()
    d = tempfile.mkdtemp(prefix="smr_cfg4_")
    try:
        extra = ["-sam", "-fastx", "-other", "-paired_in"]
        ref, _ = _run_binary("sortmerna_ref", fastas, reads, idx, os.path.join(d, "ref"), extra)
        got, log = _run_binary("sortmerna_gpu", fastas, reads, idx, os.path.join(d, "got"), extra)
        assert "Starting alignment (libsmr_b200)" in log
        assert any("passing E-value threshold = 5944" in ln for ln in got["aligned.log"]) and any("failing E-value threshold = 4056" in ln for ln in got["aligned.log"])
        assert sorted(got) == sorted(ref)
        for fn in got:
            assert got[fn] == ref[fn], fn
        # the same with the 8 indexes built on the GPU from the FASTA files (smr_build_index_device) instead of read from the index files
        dev, log = _run_binary("sortmerna_gpu", fastas, reads, idx, os.path.join(d, "dev"), extra, env=dict(os.environ, SMR_INDEX_DEVICE="1"))
        assert "Starting alignment (libsmr_b200)" in log
        for fn in ref:
            assert dev[fn] == ref[fn], "device-built index: " + fn
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_two_gpus_equal_one_gpu():
    """SMR_GPUS=2 (one context and one worker thread per GPU, batches in flight concurrently) writes what SMR_GPUS=1 writes.
    Needs two devices (gpurun --gpus 2); 2000-read batches so that both GPUs get several."""
    from sortmerna_b200 import api
    if api.load_library().smr_device_count() < 2:
        pytest.skip("needs two GPUs")
    fastas, reads, idx = _config4_inputs()
    d = tempfile.mkdtemp(prefix="smr_2gpu_")
    try:
        extra = ["-sam", "-fastx", "-other", "-paired_in"]
        one, _ = _run_binary("sortmerna_gpu", fastas, reads, idx, os.path.join(d, "g1"), extra, env=dict(os.environ, SMR_GPUS="1", SMR_BATCH_READS="2000"))
        two, log = _run_binary("sortmerna_gpu", fastas, reads, idx, os.path.join(d, "g2"), extra, env=dict(os.environ, SMR_GPUS="2", SMR_BATCH_READS="2000"))
        assert "resident on 2 GPU(s)" in log
        for fn in one:
            assert one[fn] == two[fn], fn
    finally:
        shutil.rmtree(d, ignore_errors=True)
