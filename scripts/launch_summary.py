#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`).

    python scripts/launch_summary.py x.csv > summary.txt
"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    hdr = rows[0]
    col = {k: i for i, k in enumerate(hdr)}
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        value = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        us = value / 1e3 if unit in ("ns", "nsecond") else (value * 1e3 if unit in ("ms", "msecond") else value)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + us)
        total += us
    print(f"# {sys.argv[1]}: {sum(n for n, _ in agg.values())} launches, {total / 1e3:.2f} ms of kernel time "
          f"(per-launch times under ncu are cold-cache and serialised: shares, not absolutes)")
    mine = sum(t for k, (n, t) in agg.items() if "cgb::" in k)
    print(f"# kernels of libpycgtool_b200.so: {100 * mine / total:.1f} % of the kernel time")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{100 * t / total:6.2f} %  {t / 1e3:10.3f} ms  {n:5d} x  {t / n:10.1f} us  {name[:110]}")


if __name__ == "__main__":
    main()
