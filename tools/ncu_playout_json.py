#!/usr/bin/env python3
"""profiles/r2_playout_ncu.json from an `ncu --set full` report of the playout kernel (first profiled launch)
plus the plies that launch played (taken from the bench line of the same command: outcomes.plies / steps).

    python tools/ncu_playout_json.py gpurun_out/<report>.ncu-rep <plies_per_launch> [out.json]

bench.py reads the file for the issue-slot roofline of K5: thread-instructions per ply, lanes per warp
instruction, issue-active, pipe utilisation and the kernel's real DRAM traffic -- numbers of THIS build."""
import csv
import json
import os
import subprocess
import sys


def main(path, plies, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}

    def num(name, scale_units=True):
        v, u = m[name]
        x = float(v.replace(",", ""))
        if scale_units:
            x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}.get(u, 1.0)
        return x

    warp_inst = num("smsp__inst_executed.sum")
    lanes = num("smsp__thread_inst_executed_per_inst_executed.ratio")
    res = {
        "source": os.path.basename(path) + " (ncu --set full --clock-control none, bench.py --steps 2 --warmup 3, "
                  "first timed launch)",
        "kernel": m["Kernel Name"][0] if "Kernel Name" in m else "playout_kernel",
        "plies_per_launch": plies,
        "warp_inst_per_launch": warp_inst,
        "lanes_per_warp_inst": lanes,
        "thread_inst_per_launch": warp_inst * lanes,
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "warps_per_sm": num("sm__warps_active.avg.pct_of_peak_sustained_active") * 64 / 100,
        "eligible_warps_per_cycle": num("smsp__warps_eligible.avg.per_cycle_active"),
        "registers": num("launch__registers_per_thread"),
        "duration_ms": num("gpu__time_duration.sum"),
        "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
        "dram_pct_of_peak": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    }
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]),
         sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                            "profiles", "r2_playout_ncu.json"))
