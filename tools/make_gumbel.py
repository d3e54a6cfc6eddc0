#!/usr/bin/env python
"""Records the Gumbel parameters (lambda, K) the reference computes for each bundled database with
the default scoring (Refstats::load -> 3rdparty/alp, src/sortmerna/refstats.cpp:190-233; the ALP
estimator is host-side, runs once per database and is out of scope for the GPU path).  bench.py
turns them into minimal_score for a given read set with hostio.minimal_score (refstats.cpp:236-265).

Writes sortmerna_b200/gumbel_defaults.json.  Run where oracle/_ref/sortmerna_ref and data_cache/ exist."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ora  # noqa: E402
from tools import stage_data  # noqa: E402

if __name__ == "__main__":
    stage_data.stage_inputs()
    fastas = [stage_data.db_path(n) for n in stage_data.DBS] + [os.path.join(stage_data.CACHE, "sets", "silva-bac-16s-database-id85.fasta")]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        tiny = os.path.join(d, "t.fa")
        open(tiny, "w").write(">r\n" + "ACGTTGCATGCAAGTCGAACG" * 6 + "\n")
        idx = os.path.join(stage_data.CACHE, "idx")
        r = ora.run_reference(fastas[:-1], tiny, os.path.join(d, "w"), extra=["-fastx"], idx_dir=idx)
        log = ora.parse_log(r["log"])
        for n, lam, k in zip(stage_data.DBS, log["lambda_"], log["K"]):
            out[n + ".fasta"] = dict(lambda_=lam, K=k)
        r = ora.run_reference(fastas[-1:], tiny, os.path.join(d, "w2"), extra=["-fastx"], idx_dir=os.path.join(stage_data.CACHE, "idx_set2"))
        log = ora.parse_log(r["log"])
        out["silva-bac-16s-database-id85.fasta"] = dict(lambda_=log["lambda_"][0], K=log["K"][0])
    json.dump(dict(scoring=dict(match=2, mismatch=-3, gap_open=5, gap_ext=2), gumbel=out),
              open(os.path.join(ROOT, "sortmerna_b200", "gumbel_defaults.json"), "w"), indent=1)
    print(out)
