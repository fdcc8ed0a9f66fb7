#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/*.ncu-rep, *_launches*.csv) into the small text summaries committed under profiles/.
    python profiles/summarize_ncu.py launches <launches.csv> <out.md>
    python profiles/summarize_ncu.py full <report.ncu-rep> <out.md>
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mn, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list: `{path}`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ns | share | mean ns |\n|---|---:|---:|---:|---:|\n")
        for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:90]}` | {cnt} | {ns:.0f} | {100 * ns / tot:.1f}% | {ns / cnt:.0f} |\n")
    print(open(out).read())


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full: `{path}`\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')][:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEEP:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
