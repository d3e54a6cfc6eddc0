"""The rule behind the device index builder (sortmerna_b200/csrc/smr_build_dev.cuh), checked on the CPU: a numpy model of its
steps -- distinct (L+1)-mers in order of first occurrence, per level one stable sort by (list, prefix) and the decision "a bucket
bursts at its max(17, 1 + entries present when the parent burst)-th entry" -- must reproduce, list for list and entry for entry,
what flatten_index makes of the files the REFERENCE builder wrote (tests/golden/idx), and of the host builder's files for other
options.  The GPU test (tests/test_gpu_index_build.py) then checks the kernels against the same files."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN
from sortmerna_b200 import api, hostio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAP = np.zeros(256, np.uint8)          # map_nt (indexdb.cpp:83-109)
for ch in "BCDWYbcwy":
    MAP[ord(ch)] = 1
for ch in "GKSXgksx":
    MAP[ord(ch)] = 2
for ch in "TUtu":
    MAP[ord(ch)] = 3


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("fd") / "flatten_dump")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "flatten_dump.cpp"),
                           os.path.join(ROOT, "sortmerna_b200", "csrc", "smr_index.cpp"), "-o", exe])
    return exe


def model(seqs, L=18, interval=1, max_pos=10000):
    half, pread = L // 2, L + 1
    bd = pread - half - 3
    vals, wseq, wpos = [], [], []
    for s, raw in enumerate(seqs):
        c = MAP[np.frombuffer(raw, np.uint8)].astype(np.uint64)
        nwin = (len(c) - pread + interval) // interval
        st = np.arange(nwin, dtype=np.int64) * interval
        v = np.zeros(nwin, np.uint64)
        for i in range(pread):
            v = (v << np.uint64(2)) | c[st + i]
        vals.append(v); wseq.append(np.full(nwin, s, np.uint32)); wpos.append(st.astype(np.uint32))
    v = np.concatenate(vals); wseq = np.concatenate(wseq); wpos = np.concatenate(wpos)
    n = v.size
    order = np.argsort(v, kind="stable")
    sv = v[order]
    head_e = np.concatenate(([True], sv[1:] != sv[:-1]))
    head_i = np.concatenate(([True], (sv[1:] >> np.uint64(2)) != (sv[:-1] >> np.uint64(2))))
    ids_sorted = np.cumsum(head_i) - 1
    win_id = np.empty(n, np.int64); win_id[order] = ids_sorted
    nids = int(ids_sorted[-1]) + 1
    ev, earr, eid = sv[head_e], order[head_e], ids_sorted[head_e]

    def ch(k):
        return ((ev >> np.uint64(2 * (pread - 1 - k))) & np.uint64(3)).astype(np.int64)
    kf = (ev >> np.uint64(2 * (half + 1))).astype(np.int64)
    kr = (ev & np.uint64((1 << (2 * half)) - 1)).astype(np.int64)
    tf = np.zeros(ev.size, np.int64); tr = tf.copy(); pf = tf.copy(); pr = tf.copy()
    for k in range(half + 1):
        cf, cr = ch(half + k), ch(half - k)
        tf |= cf << (2 * k); tr |= cr << (2 * k)
        if k < bd:
            pf = (pf << 2) | cf; pr = (pr << 2) | cr
    e_list = np.concatenate((kf * 2, kr * 2 + 1)); e_pref = np.concatenate((pf, pr)); e_text = np.concatenate((tf, tr))
    e_id = np.concatenate((eid, eid)); e_arr = np.concatenate((earr, earr))
    E = e_list.size
    leaf = np.zeros(E, np.int64); tpar = np.zeros(E, np.int64)
    perm = np.argsort(e_arr, kind="stable")
    for d in range(1, bd + 1):
        depth = np.where(leaf[perm] > 0, leaf[perm], d)
        drop = 2 * (bd - depth)
        key = (e_list[perm] << (2 * bd)) | ((e_pref[perm] >> drop) << drop)
        o = np.argsort(key, kind="stable")
        perm, key = perm[o], key[o]
        if d == bd:
            break
        heads = np.flatnonzero(np.concatenate(([True], key[1:] != key[:-1])))
        ends = np.concatenate((heads[1:], [E]))
        for a, b in zip(heads, ends):
            if leaf[perm[a]]:
                continue
            run = perm[a:b]
            m0 = int((e_arr[run] < tpar[run[0]]).sum())
            jb = max(17, m0 + 1)
            if b - a >= jb:
                tpar[run] = e_arr[run[jb - 1]] + 1
            else:
                leaf[run] = d
    flist = np.stack((e_text[perm], e_id[perm]), 1)
    lst = e_list[perm]
    nk = 1 << (2 * half)
    cnt = np.bincount(lst, minlength=2 * nk)
    off = np.concatenate(([0], np.cumsum(cnt)[:-1]))
    # positions
    o = np.lexsort((np.arange(n), win_id))
    wid = win_id[o]
    start = np.flatnonzero(np.concatenate(([True], wid[1:] != wid[:-1])))
    rank = np.arange(n) - np.repeat(start, np.diff(np.concatenate((start, [n]))))
    keep = (rank < max_pos) if max_pos else np.ones(n, bool)
    pos = np.stack((wpos[o][keep], wseq[o][keep]), 1)
    pcnt = np.bincount(wid[keep], minlength=nids)
    pos_off = np.concatenate(([0], np.cumsum(pcnt)))
    return dict(off=off, cnt=cnt, flist=flist, pos_off=pos_off, pos=pos)


def load_dump(dumper, prefix, part, lnwin, tmp):
    subprocess.check_call([dumper, prefix, str(part), str(lnwin), str(tmp)])
    rd = lambda n: np.fromfile(os.path.join(tmp, n), np.uint32)
    return dict(flookup=rd("flookup.u32").reshape(-1, 4), flist=rd("flist.u32").reshape(-1, 2), pos_off=rd("pos_off.u32"), pos=rd("pos.u32").reshape(-1, 2))


def compare(m, f, what):
    cnt = np.stack((f["flookup"][:, 1], f["flookup"][:, 3]), 1).reshape(-1)
    off = np.stack((f["flookup"][:, 0], f["flookup"][:, 2]), 1).reshape(-1)
    assert np.array_equal(m["cnt"], cnt), what
    assert np.array_equal(m["off"][cnt > 0], off[cnt > 0]), what
    assert np.array_equal(m["flist"][:, 0], f["flist"][:, 0]), what + ": order of the entries"
    canon = lambda po, p: (p[po[:-1].astype(np.int64), 1].astype(np.uint64) << np.uint64(32)) | p[po[:-1].astype(np.int64), 0].astype(np.uint64)
    ca, cb = canon(m["pos_off"], m["pos"]), canon(f["pos_off"], f["pos"])
    assert np.array_equal(ca[m["flist"][:, 1]], cb[f["flist"][:, 1]]), what + ": ids"
    oa, ob = np.argsort(ca), np.argsort(cb)
    assert np.array_equal(np.diff(m["pos_off"])[oa], np.diff(f["pos_off"].astype(np.int64))[ob]), what + ": position counts"


@pytest.mark.parametrize("k", [0, 1])
def test_model_equals_reference_built_index(dumper, golden, tmp_path, k):
    _, seqs, _ = hostio.read_fastx(os.path.join(GOLDEN, ("db_arc.fasta", "db_bac.fasta")[k]))
    compare(model(seqs), load_dump(dumper, golden["prefixes"][k], 0, 18, tmp_path), "golden %d" % k)


@pytest.mark.parametrize("kw", [dict(max_pos=3), dict(interval=2), dict(lnwin=16)])
def test_model_equals_host_builder_with_options(dumper, tmp_path, kw):
    fasta = os.path.join(GOLDEN, "db_bac.fasta")
    prefix = str(tmp_path / "idx")
    api.build_index(fasta, prefix, **kw)
    _, seqs, _ = hostio.read_fastx(fasta)
    L = kw.get("lnwin", 18)
    compare(model(seqs, L, kw.get("interval", 1), kw.get("max_pos", 10000)), load_dump(dumper, prefix, 0, L, tmp_path), str(kw))
