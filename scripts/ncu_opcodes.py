"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp instructions, stall samples.

usage: ncu -i x.ncu-rep --page source --csv > src.csv; python scripts/ncu_opcodes.py src.csv [top]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    inst = collections.Counter()
    thr = collections.Counter()
    samp = collections.Counter()
    stalls = collections.Counter()
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[col["Source"]].strip()
        parts = src.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        op = op.rstrip(";")
        base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDS", "STS", "LDG", "STG", "MUFU", "ATOMS", "RED", "ATOMG") else op.split(".")[0]
        n = int(r[col["Instructions Executed"]])
        inst[base] += n
        thr[base] += int(r[col["Thread Instructions Executed"]])
        samp[base] += int(r[col["# Samples"]])
        total += n
        for k in stall_cols:
            stalls[k] += int(r[col[k]] or 0)
    print(f"total warp instructions {total}")
    for op, n in inst.most_common(top):
        print(f"{op:14s} {n:12d} {100.0 * n / total:6.2f}%  lanes {thr[op] / max(n, 1):5.1f}  samples {samp[op]}")
    tot_s = sum(stalls.values())
    print("stall samples:", ", ".join(f"{k[6:]} {100.0 * v / tot_s:.1f}%" for k, v in stalls.most_common(10)))


if __name__ == "__main__":
    main()
