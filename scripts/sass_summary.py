#!/usr/bin/env python
"""SASS evidence of the shipped library (no GPU needed): per kernel, the instruction count and how many of the
Blackwell-specific mnemonics it holds -- UBLKCP (1-D bulk asynchronous copy, the TMA engine), SYNCS (mbarrier),
FFMA2 / FADD2 / FMUL2 (packed float32x2 arithmetic), MUFU, ATOMS / RED -- plus the resources ptxas reports.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
LIB = ROOT / "pycgtool_b200" / "csrc" / "libpycgtool_b200.so"
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "ATOMS", "RED", "REDG", "LDS", "STS",
         "LDG", "STG", "LDL", "STL", "SHFL", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            usage[name] = line.strip()
            name = None
    demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
    print(f"# {LIB.relative_to(ROOT)}: cuobjdump -sass / -res-usage, sm_100a")
    print("# UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier ops, FFMA2/FADD2/FMUL2 = packed float32x2")
    for block in re.split(r"\n\s*Function : ", sass)[1:]:
        fname = block.split("\n", 1)[0].strip()
        ops = collections.Counter()
        total = 0
        for line in block.split("\n"):
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                total += 1
                ops[m.group(1)] += 1
        shown = ", ".join(f"{k} {ops[k]}" for k in WATCH if ops.get(k))
        print(f"\n{demangle(fname)}\n  instructions {total}; {shown}\n  {usage.get(fname, '')}")


if __name__ == "__main__":
    main()
