#!/usr/bin/env python3
"""One line per kernel launch group from an `ncu --csv --metrics ...` log: mean duration, DRAM bytes, DRAM %, shared-pipe
wavefront %, issue-active %.  usage: ncu_api_summary.py gpurun_out/<log>.csv"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        key = (r[ix["Kernel Name"]][:90], r[ix["Metric Name"]])
        per.setdefault(key, []).append((float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]))
    kernels = collections.OrderedDict()
    for (k, m), vals in per.items():
        tail = vals[len(vals) // 2:]  # the later launches (warm)
        kernels.setdefault(k, {})[m] = (sum(v for v, _ in tail) / len(tail), tail[0][1], len(vals))
    for k, ms in kernels.items():
        print(k)
        for m, (v, u, cnt) in ms.items():
            print("    %-70s %14.4f %-8s (%d launches)" % (m, v, u, cnt))


if __name__ == "__main__":
    main(sys.argv[1])
