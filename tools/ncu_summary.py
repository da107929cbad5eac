#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel headline metrics, optionally the hottest source lines.
usage: tools/ncu_summary.py report.ncu-rep [--lines KERNEL_SUBSTR [N]]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "winst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shwave"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_shared_mem", "lim_smem"), ("launch__occupancy_limit_registers", "lim_regs")]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        print("==", d["Kernel Name"][0][:70])
        line = []
        for k, short in KEYS:
            if k in d:
                v, u = d[k]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                line.append(f"{short}={v}{u if u not in ('%', '') and short in ('time', 'dram_rd', 'dram_wr') else ''}")
        print("   " + "  ".join(line))


def lines(rep, kernel, n):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k",
                          f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, agg, stall = None, None, {}, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
            st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        elif hdr and r[0].strip().isdigit():
            try:
                s, ins = float(r[iS]), float(r[iI])
            except ValueError:
                continue
            key = (cur, int(r[0]), r[1].strip()[:90])
            a = agg.setdefault(key, [0, 0, {}])
            a[0] += s
            a[1] += ins
            for i in st:
                try:
                    a[2][hdr[i][6:]] = a[2].get(hdr[i][6:], 0) + float(r[i])
                except (ValueError, IndexError):
                    pass
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    tot = {}
    for v in agg.values():
        for k, x in v[2].items():
            tot[k] = tot.get(k, 0) + x
    print("stall mix:", {k: f"{100 * x / ts:.1f}%" for k, x in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
        top = sorted(v[2].items(), key=lambda kv: -kv[1])[:2]
        print(f"{k[0][:12]:12s} L{k[1]:4d} {100 * v[0] / ts:5.1f}%s {100 * v[1] / ti:5.1f}%i {[a for a, b in top]} {k[2]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    if "--lines" in sys.argv:
        i = sys.argv.index("--lines")
        lines(rep, sys.argv[i + 1], int(sys.argv[i + 2]) if len(sys.argv) > i + 2 else 30)
    else:
        raw(rep)
