#!/usr/bin/env python3
"""Opcode mix of one kernel in the built library (cuobjdump -sass): instruction counts per opcode class,
to see which pipe a change moved work to before spending GPU time.  usage: sass_mix.py <lib.so> <substring of the mangled kernel name>"""
import collections
import re
import subprocess
import sys


def main(lib, needle):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, mix, total = None, collections.Counter(), 0
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or needle not in cur:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            mix[m.group(2).split(".")[0]] += 1
            total += 1
    print("total", total)
    alu = sum(v for k, v in mix.items() if k in ("LOP3", "SHF", "IADD3", "ISETP", "SEL", "PRMT", "LEA", "VIADD", "POPC", "FLO", "BREV", "IABS", "VIMNMX", "PLOP3", "IMNMX", "VIADDMNMX", "LOP", "SGXT", "BMSK", "VABSDIFF4"))
    fma = sum(v for k, v in mix.items() if k in ("IMAD", "FFMA", "FMUL", "FADD"))
    print("alu-pipe ~", alu, " fma-pipe ~", fma)
    for k, v in mix.most_common(30):
        print("%-10s %d" % (k, v))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
