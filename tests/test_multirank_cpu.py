"""world_size-2 test of the N>1 plumbing on CPU (gloo): record sharding + the one all-reduce of the counter
vector.  The compute on each rank is the oracle (no GPU here); what is under test is sortmerna_b200/shard.py:
shards are disjoint and cover the batch, per-read results concatenate to the single-rank answer, and the
all-reduced counters equal the single-rank counters."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, idx_dir, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import GOLDEN, load_case
    from oracle import ora
    from sortmerna_b200 import hostio, shard
    refs = [hostio.load_references(os.path.join(GOLDEN, n)) for n in ("db_arc.fasta", "db_bac.fasta")]
    pre = hostio.find_index_prefixes(idx_dir)
    oix = [ora.OracleIndex(pre[n], 0, 18) for n in ("db_arc.fasta", "db_bac.fasta")]
    h, s, q = hostio.read_fastx(os.path.join(GOLDEN, "reads_mix.fq"))
    lo, hi = shard.shard_bounds(len(s), rank, world)
    batch = hostio.pack_reads(h[lo:hi], s[lo:hi], q[lo:hi])
    ms = load_case("default")["log"]["minimal_score"]           # refstats totals are GLOBAL, not per shard
    out = ora.align(oix, [0, 1], [0, 0], 2, refs, ms, [18, 9, 3, 18, 9, 3], ora.default_params(), batch)
    names = ("num_aligned", "num_short_last", "sw_calls", "sw_cells")
    vec = shard.allreduce_counters(shard.counter_vector(out["counters"], out["matched"], names))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=lo, hi=hi, res=out["res"], vec=vec,
             score=out["alns"]["score1"], ref=out["alns"]["ref_num"])
    dist.destroy_process_group()


def test_two_rank_sharding_and_counter_allreduce(golden, golden_idx_dir, tmp_path):
    from oracle import ora
    from sortmerna_b200 import shard
    from conftest import load_case
    world = 2
    mp.spawn(_worker, args=(world, 29731, golden_idx_dir, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    n = golden["batch"].n
    assert int(parts[0]["lo"]) == 0 and int(parts[0]["hi"]) == int(parts[1]["lo"]) and int(parts[1]["hi"]) == n
    oix = [ora.OracleIndex(p, 0, 18) for p in golden["prefixes"]]
    ms = load_case("default")["log"]["minimal_score"]
    full = ora.align(oix, [0, 1], [0, 0], 2, golden["refs"], ms, [18, 9, 3, 18, 9, 3], ora.default_params(), golden["batch"])
    assert np.array_equal(np.concatenate([p["res"] for p in parts]), full["res"])
    assert np.array_equal(np.concatenate([p["score"] for p in parts]), full["alns"]["score1"])
    names = ("num_aligned", "num_short_last", "sw_calls", "sw_cells")
    want = shard.counter_vector(full["counters"], full["matched"], names)
    for p in parts:
        assert np.array_equal(p["vec"], want)


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (10, 4), (1000001, 8)])
def test_shard_bounds_cover(n, world):
    from sortmerna_b200 import shard
    b = [shard.shard_bounds(n, r, world) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == n
    assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in b]
    assert max(sizes) - min(sizes) <= 1
