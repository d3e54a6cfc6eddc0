"""Run a sortmerna host binary (the reference, or the reference linked with integration/align_gpu.cpp) and normalise its output
files for comparison."""
import os
import re
import subprocess

from conftest import GOLDEN, ROOT

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_VOLATILE = re.compile(r"Time|time|Command|Process pid|/tmp/|Date|sec|\d\d:\d\d:\d\d")


def run_host(binary, workdir, reads, extra, threads=2):
    cmd = [os.path.join(REF_DIR, binary), "-ref", os.path.join(GOLDEN, "db_arc.fasta"), "-ref", os.path.join(GOLDEN, "db_bac.fasta")]
    for r in reads:
        cmd += ["-reads", r]
    cmd += ["-workdir", workdir, "-threads", str(threads), "-task", "4"] + list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    out = {}
    d = os.path.join(workdir, "out")
    for fn in sorted(os.listdir(d)):
        lines = open(os.path.join(d, fn), errors="replace").read().split("\n")
        if fn.endswith(".log"):
            lines = [ln for ln in lines if not _VOLATILE.search(ln)]
        elif fn.endswith(".sam"):
            lines = [ln for ln in lines if not ln.startswith("@PG")]       # carries the command line
        out[fn] = lines
    return out, p.stdout


_PER_DB = re.compile(r"^(\s+\S+\.fasta\t+)([0-9.]+)\s*$")


def assert_same_outputs(a, b, ref_threads=2):
    """a: our run, b: the reference binary's.  With several threads the reference's "Coverage by database" figures come from
    ++readstats.reads_matched_per_db[...] without synchronisation (src/sortmerna/alignment.cpp:415,454-457): now and
    then a run loses an update (66.33 % instead of 66.67 % of 300 reads).  Those lines are compared with a tolerance of one
    percentage point then, and exactly when the reference ran single-threaded (ref_threads=1); everything else is always exact."""
    assert sorted(a) == sorted(b), (sorted(a), sorted(b))
    for fn in a:
        x, y = list(a[fn]), list(b[fn])
        if ref_threads > 1 and fn.endswith(".log") and len(x) == len(y):
            for i, (p, q) in enumerate(zip(x, y)):
                mp, mq = _PER_DB.match(p), _PER_DB.match(q)
                if mp and mq and mp.group(1) == mq.group(1) and abs(float(mp.group(2)) - float(mq.group(2))) <= 1.0:
                    y[i] = p
        assert x == y, f"{fn} differs: first difference {next((p, q) for p, q in zip(x, y) if p != q)}"
