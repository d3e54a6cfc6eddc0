"""Random (database slice, reads, option set) generator shared by the differential fuzz tests."""
import os

import numpy as np

from conftest import GOLDEN
from sortmerna_b200 import hostio


def make_case(seed: int, workdir: str):
    rng = np.random.default_rng(77000 + seed)
    fastas = []
    for name in ("db_arc.fasta", "db_bac.fasta"):
        h, s, _ = hostio.read_fastx(os.path.join(GOLDEN, name))
        pick = sorted(rng.choice(len(h), size=int(rng.integers(15, 50)), replace=False).tolist())
        p = os.path.join(workdir, f"fz{seed}_{name}")
        with open(p, "w") as f:
            for k in pick:
                f.write(h[k] + "\n" + s[k].decode() + "\n")
        fastas.append(p)
    h, s, q = hostio.read_fastx(os.path.join(GOLDEN, "reads_mix.fq"))
    pick = rng.choice(len(h), size=250, replace=False)
    reads = []
    for n, k in enumerate(pick):
        seq = list(s[k].decode())
        if rng.random() < 0.4 and len(seq) > 30:       # extra damage: substitutions, an indel, an N
            for _ in range(int(rng.integers(1, 6))):
                seq[int(rng.integers(len(seq)))] = "ACGTN"[int(rng.integers(5))]
            if rng.random() < 0.5:
                del seq[int(rng.integers(len(seq)))]
            else:
                seq.insert(int(rng.integers(len(seq))), "ACGT"[int(rng.integers(4))])
        reads.append((f"fz{n}", "".join(seq)))
    reads_p = os.path.join(workdir, f"fz{seed}_reads.fq")
    with open(reads_p, "w") as f:
        for name, sq in reads:
            f.write(f"@{name}\n{sq}\n+\n{'I' * len(sq)}\n")
    # option set (reference CLI arguments); scoring sets restricted to ones the reference's ALP tables accept
    scoring = [[], ["-match", "2", "-mismatch", "-4", "-gap_open", "6", "-gap_ext", "3", "-N", "-2"],
               ["-match", "2", "-mismatch", "-7", "-gap_open", "3", "-gap_ext", "1"], ["-match", "1", "-mismatch", "-2", "-gap_open", "3", "-gap_ext", "2"]]
    args = list(scoring[int(rng.integers(len(scoring)))])
    na = int(rng.choice([1, 1, 2, 4]))
    if na != 1:
        args += ["-num_alignments", str(na)]
    if rng.random() < 0.3:
        args += ["-no-best"]
    r = rng.random()
    if r < 0.15:
        args += ["-F"]
    elif r < 0.3:
        args += ["-R"]
    if rng.random() < 0.2:
        args += ["-full_search"]
    if rng.random() < 0.4:
        args += ["-num_seeds", str(int(rng.integers(1, 4)))]
    if rng.random() < 0.4:
        args += ["-min_lis", str(int(rng.integers(1, 4)))]
    if rng.random() < 0.4:
        args += ["-edges", str(int(rng.integers(1, 9))) + ("%" if rng.random() < 0.3 else "")]
    return fastas, reads_p, args
