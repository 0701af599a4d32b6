#!/usr/bin/env python3
"""Stall density along the hot loop of a kernel in an ncu report: consecutive windows of SASS instructions
with their share of all stall samples, the dominant stall reasons and the opcode mix."""
import csv
import subprocess
import sys


def main(path, min_exec_frac=0.5, win=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    top_exec = max(float(r[ix["Instructions Executed"]] or 0) for r in data)
    hot = [(i, r) for i, r in enumerate(data) if float(r[ix["Instructions Executed"]] or 0) > min_exec_frac * top_exec]
    base = sum(int(r[ix["stall_selected"]] or 0) for _, r in hot) / max(1, len(hot))
    print("instructions %d, hot %d, samples %d, selected/instr %.0f" % (len(data), len(hot), tot, base))
    for j in range(0, len(hot), win):
        seg = hot[j:j + win]
        s = sum(int(r[ix["# Samples"]]) for _, r in seg)
        agg = {h: sum(int(r[ix[h]] or 0) for _, r in seg) for h in stalls}
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:3]
        ops = {}
        for _, r in seg:
            toks = r[ix["Source"]].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
        print("%4d-%4d %5.1f%% density=%4.1f thr=%4.1f %s | %s" % (
            seg[0][0], seg[-1][0], 100 * s / tot, s / len(seg) / base,
            sum(float(r[ix["Avg. Threads Executed"]] or 0) for _, r in seg) / len(seg),
            [(k.replace("stall_", ""), v) for k, v in top], sorted(ops.items(), key=lambda kv: -kv[1])[:5]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.2)
