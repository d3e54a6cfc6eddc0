"""The seed kernels replace the reference's table-driven Levenshtein automaton (traverse_bursttrie.cpp:68-98)
by bit-parallel edit-distance predicates (sortmerna_b200/csrc/smr_levbits.h).  This host-side check proves,
on millions of random (half-window, trie text) pairs, that table == edit distance <= 1 == bit formulas."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_automaton_equals_edit_distance_equals_bit_formulas():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "lev_bits_check")
        subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "lev_bits_check.cpp"), "-o", exe])
        out = subprocess.run([exe, "1500000"], stdout=subprocess.PIPE, text=True)
        assert out.returncode == 0, out.stdout
        f = out.stdout.split()
        assert f[0] == "cases" and [f[i] for i in (3, 5, 7, 9)] == ["0", "0", "0", "0"], out.stdout
