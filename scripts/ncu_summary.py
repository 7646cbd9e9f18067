#!/usr/bin/env python3
"""Summarise an .ncu-rep (run where ncu is installed): key raw metrics + stall breakdown + hottest SASS lines.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for vals in raw[2:]:
        print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {vals[i]} {units[i]}")
        for i, h in enumerate(hdr):
            if "per_issue_active" in h and vals[i] and float(vals[i]) > 0.05:
                print(f"{h:75s} {vals[i]}")
    src = page(rep, "source")
    hdr, data = src[1], src[2:]
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    isamp, iex, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[isamp] or 0) for r in data) or 1
    agg = {n: sum(int(r[hdr.index(n)] or 0) for r in data) / tot for n in names}
    print("== warp stall sampling shares (all samples):")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        if v > 0.002:
            print(f"   {n:28s} {v:.3f}")
    print("== hottest SASS instructions (share of samples, executions, instruction):")
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:16]:
        print(f"   {int(r[isamp] or 0) / tot:6.3f} {r[iex]:>12s}  {r[isrc][:90]}")
    sass = " ".join(r[isrc] for r in data)
    for m in ("UBLKCP", "SYNCS", "LDGSTS", "LDG.E", "LOP3", "UTMALDG"):
        print(f"== SASS has {m}: {m in sass}")


if __name__ == "__main__":
    main()
