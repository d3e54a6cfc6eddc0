"""The kernels against the oracle on the random cases of tests/fuzz_common.py (database slices, damaged reads, random option
sets): identical per-read state, alignments, CIGARs, report arithmetic."""
import os
import shutil
import tempfile

import numpy as np
import pytest

from fuzz_common import make_case
from helpers import assert_same_results, params_kwargs_from_args
from sortmerna_b200 import api, hostio

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(10))
def test_kernels_equal_oracle_on_random_cases(seed):
    from oracle import ora
    d = tempfile.mkdtemp(prefix="smr_fzg_")
    try:
        fastas, reads_p, args = make_case(seed, d)
        rng = np.random.default_rng(seed)
        ms = [int(rng.integers(25, 70)), int(rng.integers(25, 70))]      # any threshold will do for kernel-vs-oracle
        batch = hostio.load_reads(reads_p)
        refs = [hostio.load_references(f) for f in fastas]
        kw = params_kwargs_from_args(args)
        al = api.Aligner(0)
        al.set_params(api.default_params(**kw))
        oix = []
        for k, f in enumerate(fastas):
            p = os.path.join(d, f"idx{k}")
            api.build_index(f, p)
            al.load_index_part(k, 0, p, refs[k], ms[k], (18, 9, 3), 18)
            oix.append(ora.OracleIndex(p, 0, 18))
        got = al.align(batch.cat, batch.off, with_stats=True)
        want = ora.align(oix, [0, 1], [0, 0], 2, refs, ms, [18, 9, 3, 18, 9, 3], ora.default_params(**kw), batch, nthreads=4)
        assert_same_results(got, want, str(args))
        assert got["matched"].tolist() == want["matched"].tolist()
        st = hostio.host_aln_stats(batch, refs, got["res"], got["alns"], got["cigar"], got["slots"])
        live = np.zeros(batch.n * got["slots"], bool)
        for r in range(batch.n):
            live[r * got["slots"]:r * got["slots"] + int(got["res"]["n_align"][r])] = True
        for f in ("n_miss", "n_gap", "n_match", "n_match_denovo"):
            assert np.array_equal(got["stats"][f][live], st[f][live]), (f, args)
        al.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)
