#!/usr/bin/env python
"""Turn the files one `tools/gpu_evidence.sh <tag>` call left in gpurun_out/ into the tracked evidence under profiles/
(read here, no GPU needed):  bench lines, launch-share table, --set full summary + hottest lines, traffic.json (stamped
with the current commit), SASS listing of the hot kernels of the library as built.
usage: tools/profiles_from_evidence.py <tag in gpurun_out> <tag in profiles>"""
import collections
import csv
import gzip
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ("interp_qgpv_kernel", "lwa_window_kernel", "lwa_prefix_kernel", "flux_column_kernel", "area_hist3_kernel", "zonal_stats_kernel",
       "theta_rowmean_kernel")
CMD = "python bench.py --kernels-only --steps 1 --warmup 1 --batch 1"


def short(name):
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0]


def launch_shares(src, dst):
    rows = [r for r in csv.DictReader(l for l in open(src) if l.startswith('"'))]
    tot = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r["Metric Value"]) / 1e3
    whole = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list of `{CMD}` (gpu__time_duration.sum, --clock-control none): every launch of the\n"
                "process (plan creation, two warm-up/timed steps of one 0.25-degree snapshot, the profiled step). "
                "Cold-cache, serialised: compare SHARES.\n\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:60s} launches {n:3d}  total {us:9.1f} us  share {100 * us / whole:5.1f}%\n")


def full_summary(rep, dst):
    tool = os.path.join(ROOT, "tools", "ncu_summary.py")
    with open(dst, "w") as f:
        f.write(subprocess.run([sys.executable, tool, rep], capture_output=True, text=True).stdout)
        for k in ("lwa_window", "interp_qgpv", "flux_column", "area_hist3"):
            f.write(f"---- hottest source lines: {k}\n")
            f.write(subprocess.run([sys.executable, tool, rep, "--lines", k, "12"], capture_output=True, text=True).stdout)


def traffic(rep, dst, commit):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = collections.OrderedDict()
    res["_commit"] = commit
    res["_how"] = f"ncu --set full --clock-control none, {CMD[7:]} (one 0.25-degree snapshot), tools/gpu_evidence.sh"
    res["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one snapshot"
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = short(d["Kernel Name"])
        key = name.split("<")[0]
        if key == "flux_column_kernel":
            key = "flux_special_rows" if name.startswith("flux_column_kernel<1") else "flux_uvt_column_kernel"
        b = sum(float(d[m].replace(",", "")) * scale[u[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[u["gpu__time_duration.sum"]]
        e = res.setdefault(key, {"dram_bytes_per_snapshot": 0.0, "launches": 0, "time_us": 0.0})
        e["dram_bytes_per_snapshot"] += b
        e["launches"] += 1
        e["time_us"] = round(e["time_us"] + t, 3)
    res["_total_dram_bytes_per_snapshot"] = sum(v["dram_bytes_per_snapshot"] for k, v in res.items() if not k.startswith("_"))
    json.dump(res, open(dst, "w"), indent=1)
    return res


def sass(dst_gz, dst_summary):
    lib = os.path.join(ROOT, "hn2016_falwa_b200", "libfalwa_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    parts = out.split("Function : ")
    keep, summ = [], []
    for p in parts[1:]:
        mangled = p.split("\n", 1)[0].strip()
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        if not any(h in name for h in HOT):
            continue
        keep.append("Function : " + name + "\n" + p)
        ops = collections.Counter(m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)[\. ;]", p))
        sel = {k: ops[k] for k in ("UTMALDG", "LDTM", "STTM", "SYNCS", "UTCBAR", "UTCHMMA", "HMMA", "DMMA", "DFMA", "DADD", "DMUL",
                                   "DSETP", "LDG", "STG", "LDS", "STS", "ATOMS", "SHFL", "VOTE", "BAR", "MUFU", "CCTL") if ops[k]}
        summ.append(f"{short(name)[:70]:70s} {sum(ops.values()):6d} instr  {sel}")
    with gzip.open(dst_gz, "wt") as f:
        f.write("\n".join(keep))
    with open(dst_summary, "w") as f:
        f.write("SASS of the hot kernels in hn2016_falwa_b200/libfalwa_b200.so (cuobjdump -sass; full text in the .gz next to this "
                "file).\nUTMALDG = TMA tensor load, LDTM/STTM = tensor-memory load/store, SYNCS = mbarrier; no HMMA/DMMA/UTCHMMA "
                "(tensor-core math) anywhere.\n\n" + "\n".join(sorted(summ)) + "\n")


def main(src, dst):
    G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    for s in ("bench.json", "bench_upsampled.json", "bench_reference.json", "config4.json", "config5.json", "fp64_peak.json", "membw.txt",
              "ncu_launches.csv"):
        a = os.path.join(G, f"{src}_{s}")
        if os.path.exists(a):
            shutil.copy(a, os.path.join(P, f"{dst}_{s}"))
    launch_shares(os.path.join(G, f"{src}_ncu_launches.csv"), os.path.join(P, f"{dst}_ncu_launch_shares.txt"))
    rep = os.path.join(G, f"{src}_prof.ncu-rep")
    full_summary(rep, os.path.join(P, f"{dst}_ncu_full_summary.txt"))
    t = traffic(rep, os.path.join(P, "traffic.json"), commit)
    sass(os.path.join(P, f"{dst}_sass_hot_kernels.txt.gz"), os.path.join(P, f"{dst}_sass_summary.txt"))
    if os.path.exists(os.path.join(G, "parity_fullsize.jsonl")):
        shutil.copy(os.path.join(G, "parity_fullsize.jsonl"), os.path.join(P, f"{dst}_parity_fullsize.jsonl"))
    print("commit", commit, "DRAM GB/snapshot", t["_total_dram_bytes_per_snapshot"] / 1e9)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
