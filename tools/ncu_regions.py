#!/usr/bin/env python3
"""Summarise an ncu report's source page: SASS regions grouped by execution count, with the share of all
warp instructions and the average number of active threads (divergence)."""
import csv
import subprocess
import sys


def main(path, min_share=0.004):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
    thr = sum(float(r[ix["Thread Instructions Executed"]] or 0) for r in data)
    print("kernel:", rows[0][1])
    print("warp instructions %.4g, thread instructions %.4g, avg active threads %.2f" % (tot, thr, thr / tot))
    groups, cur = [], None
    for r in data:
        ie = float(r[ix["Instructions Executed"]] or 0)
        at = float(r[ix["Avg. Threads Executed"]] or 0)
        sm = int(r[ix["# Samples"]] or 0)
        if cur and abs(cur["ie"] - ie) < 0.02 * max(cur["ie"], 1):
            cur["n"] += 1; cur["tot"] += ie; cur["samples"] += sm; cur["thr"] += ie * at
        else:
            cur = {"ie": ie, "n": 1, "tot": ie, "first": r[ix["Source"]], "samples": sm, "thr": ie * at,
                   "addr": r[ix["Address"]]}
            groups.append(cur)
    for g in groups:
        if g["tot"] / tot > min_share:
            print("%6s n=%3d exec/instr=%8.2fM share=%5.1f%% avgthr=%5.1f samples=%7d | %s" % (
                g["addr"][-5:], g["n"], g["ie"] / 1e6, 100 * g["tot"] / tot, g["thr"] / max(g["tot"], 1),
                g["samples"], g["first"][:50]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.004)
