"""Parity at the size of the reference's own integration tests (BASELINE configs 2 and 4), on the bundled read sets:
the GPU path against what the unmodified reference binary prints ON THE SAME BOX for the same inputs -- every SAM row
and the pass/fail totals -- plus the known-answer counts of scripts/test.jinja (t3: 99999/1, t17: 5944/4056).
Needs data_cache/ (tools/stage_data.py; travels with the gpurun snapshot) and oracle/_ref/sortmerna_ref."""
import os
import tempfile

import numpy as np
import pytest

from sortmerna_b200 import api, hostio

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.path.join(ROOT, "data_cache")


def _need(*paths):
    from oracle import ora
    if not ora.have_reference_binary():
        pytest.skip("oracle/_ref/sortmerna_ref not built")
    for p in paths:
        if not os.path.exists(p):
            pytest.skip(f"{p} not staged (tools/stage_data.py)")


def _run_case(fastas, idx_dir, read_files, ref_extra, threads, **native_kw):
    """The reference runs on the index ITS OWN builder makes (<idx_dir>_ref); the GPU path runs on the index smr_build_index
    makes (<idx_dir>) -- so the comparison also covers the native index builder at full database size."""
    from oracle import ora
    from tools import stage_data
    ref_idx, _ = stage_data.ensure_indexes(fastas, idx_dir + "_ref", extra=tuple(ref_extra), builder="reference")   # index-build options (-max_pos) == run options here
    idx_dir, _ = stage_data.ensure_indexes(fastas, idx_dir, **native_kw)
    with tempfile.TemporaryDirectory(prefix="smr_sets_") as d:
        r = ora.run_reference(fastas, read_files, os.path.join(d, "w"), extra=["-sam", "-fastx", "-other"] + list(ref_extra), threads=threads, idx_dir=ref_idx)
        log = ora.parse_log(r["log"])
        sam = ora.read_sam_rows(os.path.join(r["out_dir"], "aligned.sam"))
    pre = hostio.find_index_prefixes(idx_dir)
    refs = [hostio.load_references(f) for f in fastas]
    al = api.Aligner(0)
    al.set_params(api.default_params())
    for k, f in enumerate(fastas):
        p = pre[os.path.basename(f)]
        al.load_index_part(k, 0, p, refs[k], log["minimal_score"][k], (18, 9, 3), hostio.parse_stats(p).lnwin)
    h, s, q = [], [], []
    for rf in read_files:
        hh, ss, qq = hostio.read_fastx(rf)
        h += hh; s += ss; q += qq
    batch = hostio.pack_reads(h, s, q)
    got = al.align(batch.cat, batch.off)
    rows = hostio.format_sam_rows(batch, refs, got["res"], got["alns"], got["cigar"], got["slots"])
    al.close()
    return log, sam, rows, got, batch


def test_set4_paired_vs_8_databases():
    """BASELINE config 4 / t17: 2 x 5000 mates vs the 8 rRNA databases."""
    from tools import stage_data
    reads = [os.path.join(CACHE, "sets", f"set4_mate_pairs_metatranscriptomics_{k}.fastq") for k in (1, 2)]
    fastas = [stage_data.db_path(n) for n in stage_data.DBS]
    _need(*reads, *fastas)
    log, sam, rows, got, batch = _run_case(fastas, os.path.join(CACHE, "idx"), reads, [], threads=os.cpu_count() or 8)
    assert (log["passing"], log["failing"]) == (5944, 4056)          # scripts/test.jinja:1186-1188
    assert int(got["res"]["is_hit"].sum()) == 5944
    assert sorted(rows) == sorted(sam)


def test_set2_amplicon_vs_bac16s_id85():
    """BASELINE config 2 / t3: 100,000 amplicon reads vs silva-bac-16s-database-id85 built with -max_pos 250."""
    reads = [os.path.join(CACHE, "sets", "set2_environmental_study_550_amplicon.fasta")]
    fastas = [os.path.join(CACHE, "sets", "silva-bac-16s-database-id85.fasta")]
    _need(*reads, *fastas)
    log, sam, rows, got, batch = _run_case(fastas, os.path.join(CACHE, "idx_set2"), reads, ["-max_pos", "250"], threads=os.cpu_count() or 8, max_pos=250)
    assert (log["passing"], log["failing"]) == (99999, 1)            # scripts/t3.jinja:30-32
    assert int(got["res"]["is_hit"].sum()) == 99999
    assert sorted(rows) == sorted(sam)


def test_bench_workload_sample_vs_reference():
    """The synthetic workload bench.py measures (BASELINE configs 3/5: 150 nt reads at 1 % / 10 % error + random, vs the 8
    databases): 20,000 reads from bench.gen_reads through the GPU path and through the unmodified reference ON THE SAME BOX --
    every SAM row and the pass/fail totals identical."""
    import bench
    from oracle import ora
    from tools import stage_data
    fastas = [stage_data.db_path(n) for n in stage_data.DBS]
    _need(*fastas)
    refs = [hostio.load_references(f) for f in fastas]
    reads = bench.gen_reads(bench.DbPool(refs), 20000, bench.GEN_SEED + 4242)
    with tempfile.TemporaryDirectory(prefix="smr_wl_") as d:
        fq = os.path.join(d, "wl.fq")
        bench.write_fastq(fq, reads)
        log, sam, rows, got, batch = _run_case(fastas, os.path.join(CACHE, "idx"), [fq], [], threads=os.cpu_count() or 8)
    assert log["passing"] + log["failing"] == 20000
    assert int(got["res"]["is_hit"].sum()) == log["passing"]
    assert 0.5 < log["passing"] / 20000 < 0.8          # two thirds of the workload derive from the databases
    assert sorted(rows) == sorted(sam)
