"""The gzip inflate of sortmerna_b200/csrc/smr_inflate.h run on the CPU (tests/inflate_check.cpp drives the same FIND / COUNT /
WRITE / WINDOW / RESOLVE steps the kernels perform, serially) against zlib -- the logic the GPU tests then only have to repeat."""
import os
import subprocess

import pytest

import inflate_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("inflate")
    e = str(d / "inflate_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "inflate_check.cpp"), "-o", e])
    return e


def run(exe, tmp_path, gz, chunk):
    src, dst = tmp_path / "in.gz", tmp_path / "out.bin"
    src.write_bytes(gz)
    p = subprocess.run([exe, str(src), str(dst), str(chunk)], capture_output=True, text=True)
    return p.returncode, p.stdout.strip(), (dst.read_bytes() if p.returncode == 0 else b"")


def test_inflate_equals_zlib(exe, tmp_path):
    spans_seen = 0
    for name, gz, want in inflate_cases.cases(3000):
        for chunk in (1 << 30, 16384, 4096):
            rc, msg, got = run(exe, tmp_path, gz, chunk)
            assert rc == 0, (name, chunk, msg)
            assert got == want, (name, chunk, msg)
            spans_seen = max(spans_seen, int(msg.split()[4]))
    assert spans_seen > 20      # the speculative block search did split the streams


def test_bad_input_is_refused(exe, tmp_path):
    for name, gz in inflate_cases.bad_cases():
        rc, msg, _ = run(exe, tmp_path, gz, 4096)
        assert rc == 1 and msg.startswith("error"), (name, msg)
