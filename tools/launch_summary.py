#!/usr/bin/env python
"""Per-kernel totals and shares of an ncu launch list (tools/profile.sh: gpu__time_duration.sum of every launch of our kernels in a
short bench run).  Per-launch times under ncu are cold-cache and serialised: the SHARES are what is compared with bench.py's
CUDA-event split.  Also prints the seed kernel's time per index part (8 launches per step, in --ref order)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt, seed = collections.Counter(), collections.Counter(), []
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    ns = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ix["Metric Unit"]], 1.0)
    tot[name] += ns; cnt[name] += 1
    if "seed_kernel" in name:
        seed.append(ns)
allns = sum(tot.values())
print(f"# ncu launch list ({sys.argv[1]}): bench.py --reads 1000000 --steps 2 --warmup 1 --no-cpu-baseline; gpu__time_duration.sum, --clock-control none")
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event split (kernel_ms_per_step)\n")
for k, v in tot.most_common():
    print(f"{k:52s} launches {cnt[k]:4d} total {v / 1e6:10.2f} ms share {100 * v / allns:5.1f}%")
nparts = 8
if seed and len(seed) % nparts == 0:
    per = [sum(seed[i::nparts]) / (len(seed) // nparts) / 1e6 for i in range(nparts)]
    print("\nseed_kernel per index part (ms per launch, mean): " + ", ".join(f"{x:.2f}" for x in per))
