"""Summarise an .ncu-rep (raw + source pages) into a short text report.  usage: ncu_summary.py file.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h:85s} {u:16s} {v}")
    rows = page(rep, "source")
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = []
    tot = tots = 0
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: 0 for h in stall}
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[ci["# Samples"]].isdigit():
            continue
        n, s = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])
        data.append((n, s, r))
        tot += n
        tots += s
        for h in stall:
            agg[h] += int(r[ci[h]])
    print(f"\nwarp instructions executed {tot}, samples {tots}")
    print("stall samples:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    print(f"\ntop {top} instructions by samples:")
    for i, (n, s, r) in sorted(enumerate(data), key=lambda x: -x[1][1])[:top]:
        st = sorted([(int(r[ci[h]]), h[6:]) for h in stall], reverse=True)[:2]
        print(f"{i:5d} exec {n:10d} ({100 * n / tot:4.1f}%) smp {100 * s / tots:5.1f}% thr {r[ci['Avg. Threads Executed']]:>3s} "
              f"{r[ci['Source']].strip()[:60]:60s} {st}")


if __name__ == "__main__":
    main()
