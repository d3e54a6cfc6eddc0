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
