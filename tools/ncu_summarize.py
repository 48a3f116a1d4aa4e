#!/usr/bin/env python
"""Reduce an `ncu --set full` report to what the repo commits under profiles/.

    python tools/ncu_summarize.py gpurun_out/prof.ncu-rep --tag r02 --steps 1 --command "<the profiled command>"

Writes
  profiles/<tag>_ncu_full_summary.csv   one row per profiled launch with the columns the roofline discussion uses
  profiles/ncu_traffic.json             per-kernel DRAM bytes per benchmark step + the hash of the kernel sources the
                                        capture was taken from; bench.py reports roofline.traffic from this file only
                                        when that hash equals the hash of the sources it is running
Runs here (no GPU needed): it only reads the report through `ncu -i`.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12,
         "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def short_name(full):
    name = full.split("(")[0].replace("void ", "")
    if "::" in name:
        name = name.split("::")[-1]
    return name.split("<")[0].strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--steps", type=int, default=1, help="benchmark steps covered by the capture")
    ap.add_argument("--command", default="")
    ap.add_argument("--no-traffic", action="store_true", help="only write the summary CSV")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {c: hdr.index(c) for c in COLS if c in hdr}
    out_csv = os.path.join(REPO, "profiles", f"{args.tag}_ncu_full_summary.csv")
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(list(idx))
        w.writerow([units[i] for i in idx.values()])
        for r in data:
            w.writerow([r[i] for i in idx.values()])
    print("wrote", out_csv, len(data), "launches")
    if args.no_traffic:
        return
    from bench import kernel_source_hash
    kernels = {}
    ir, iw, it, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name")
    for r in data:
        k = kernels.setdefault(short_name(r[ik]), {"launches": 0, "dram_bytes": 0.0, "ms_under_ncu": 0.0})
        k["launches"] += 1
        k["dram_bytes"] += float(r[ir].replace(",", "")) * SCALE.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * SCALE.get(units[iw], 1.0)
        k["ms_under_ncu"] += float(r[it].replace(",", "")) * SCALE.get(units[it], 1.0) * 1e3
    for k in kernels.values():
        k["dram_bytes_per_step"] = k["dram_bytes"] / args.steps
    out = {"source_hash": kernel_source_hash(), "command": args.command, "steps": args.steps, "kernels": kernels,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per kernel name, summed over its launches; durations are "
                   "cold-cache and serialised (ncu replay): compare shares, not absolutes"}
    path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, {k: round(v["dram_bytes_per_step"] / 1e9, 3) for k, v in kernels.items()})


if __name__ == "__main__":
    main()
