#!/usr/bin/env python
"""profiles/<round>_traffic.json from the raw metric pages tools/profile.sh exported (gpurun_out/<round>_{lis,seed}_raw.csv):
DRAM bytes, duration and pipe utilisation of the ONE captured launch of each kernel (bench.py scales the bytes to its launch size)."""
import csv
import json
import sys


SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}   # -> bytes, milliseconds


def last_values(path):
    """metric -> value of the last row, byte counts in bytes and durations in ms whatever unit ncu chose for the page"""
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, vals = rows[hdr], rows[hdr + 1], rows[-1]
    out = {}
    for n, u, v in zip(names, units, vals):
        try:
            out[n] = float(str(v).replace(",", "")) * SCALE.get(u, 1.0)
        except ValueError:
            out[n] = v
    return out


def num(v):
    return float(v)


def main(rnd, reads):
    out = {"note": f"dram__bytes_read.sum + dram__bytes_write.sum from ONE ncu --set full capture per kernel (gpurun, bench.py --reads {reads} --steps 1 --warmup 1; "
                   f"profiles/{rnd}_*_kernel_ncu_details.txt, {rnd}_*_raw.csv); bench.py scales them to its launch size (bytes are proportional to the reads in a launch)",
           "reads_in_captured_launch": reads}
    for k, kern in (("seed", "seed_kernel"), ("lis", "lis_kernel")):
        v = last_values(f"gpurun_out/{rnd}_{k}_raw.csv")
        rd, wr = num(v["dram__bytes_read.sum"]), num(v["dram__bytes_write.sum"])
        out[kern] = {"dram_bytes": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                     "duration_ms": num(v["gpu__time_duration.sum"]),
                     "alu_pipe_pct": num(v["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]),
                     "fma_pipe_pct": num(v["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]),
                     "lsu_pipe_pct": num(v["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]),
                     "l2_hit_pct": num(v["lts__t_sector_hit_rate.pct"]), "issue_active_pct": num(v["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                     "threads_per_inst": num(v["smsp__thread_inst_executed_per_inst_executed.ratio"])}
    json.dump(out, open(f"profiles/{rnd}_traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 400000)
