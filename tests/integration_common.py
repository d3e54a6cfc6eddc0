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


def assert_same_outputs(a, b):
    assert sorted(a) == sorted(b), (sorted(a), sorted(b))
    for fn in a:
        assert a[fn] == b[fn], f"{fn} differs: first difference {next((x, y) for x, y in zip(a[fn], b[fn]) if x != y)}"
