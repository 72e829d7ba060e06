#!/usr/bin/env python
"""Static view of a kernel's loops: `cuobjdump -sass` of an object / library, one function, every backward branch
as a loop with its instruction count and opcode mix (no GPU needed).

    python scripts/sass_loops.py OBJ FUNCTION_SUBSTRING [min_len]
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", text)
    body = next((f for f in funcs[1:] if pat in f.split("\n", 1)[0]), None)
    if body is None:
        sys.exit("function not found")
    print(body.split("\n", 1)[0])
    inst = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            inst.append((int(m.group(1), 16), m.group(2).strip()))
    print("instructions:", len(inst))
    addr_index = {a: i for i, (a, _) in enumerate(inst)}
    loops = []
    for i, (a, txt) in enumerate(inst):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", txt)
        if m:
            t = int(m.group(1), 16)
            if t <= a and t in addr_index:
                loops.append((addr_index[t], i))
    for lo, hi in sorted(loops, key=lambda x: x[0] - x[1]):
        n = hi - lo + 1
        if n < min_len:
            continue
        mix = collections.Counter()
        for _, txt in inst[lo:hi + 1]:
            parts = txt.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            f = op.split(".")
            key = f[0] + ("." + f[1] if f[0] in ("LDS", "STS", "MUFU", "STG", "LDG", "ATOMS", "RED") and len(f) > 1 else "")
            mix[key] += 1
        print(f"loop {inst[lo][0]:#x}..{inst[hi][0]:#x}: {n} instructions")
        print("   " + ", ".join(f"{k} {v}" for k, v in mix.most_common(40)))


if __name__ == "__main__":
    main()
