#!/usr/bin/env python
"""Per-source-line view of an ncu report (`ncu --set full --import-source on`): stall-sample totals by reason, the hottest source lines
with their dominant stall reasons, and every CALL site with a non-zero executed count (out-of-line helpers -- e.g. the slow path of an
IEEE division -- that actually run).
    python profiles/ncu_source_hotspots.py <report.ncu-rep> [top_n] > profiles/<name>_hotspots.md"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, topn):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur = hdr = line = src = None
    agg, calls = collections.OrderedDict(), []
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None:
            continue
        num = lambda v: int(v) if v.isdigit() else 0
        if r[0] != "":
            line, src = r[0], r[1][:90]
            a = agg.setdefault((cur, line), [0, 0, src, collections.Counter()])
            a[0] += num(r[i_inst]); a[1] += num(r[i_samp])
            for i, h in stall:
                a[3][h] += num(r[i])
        elif "CALL" in r[3] and num(r[i_inst]) > 0:
            calls.append((cur, line, r[3].split()[0] + " ...", num(r[i_inst]), src))
    tot_i, tot_s = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    allst = collections.Counter()
    for a in agg.values():
        allst.update(a[3])
    print(f"# ncu source hot spots: `{rep}`\n\nwarp-instructions executed: {tot_i}; stall samples: {tot_s}\n")
    print("stall reasons (% of samples): " + ", ".join(f"{k} {100 * v / max(tot_s, 1):.1f}" for k, v in allst.most_common(10)) + "\n")
    print("| file:line | instr % | samples % | top stall reasons | source |\n|---|---:|---:|---|---|")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        st = ", ".join(f"{k}:{v}" for k, v in a[3].most_common(3))
        print(f"| {f}:{l} | {100 * a[0] / max(tot_i, 1):.2f} | {100 * a[1] / max(tot_s, 1):.2f} | {st} | `{a[2].replace('|', '/')}` |")
    print("\nexecuted CALL sites (out-of-line helpers that ran):\n\n| file:line | executed (warp-level) | source |\n|---|---:|---|")
    for f, l, op, n, s_ in calls:
        print(f"| {f}:{l} | {n} | `{s_.replace('|', '/')}` |")
    if not calls:
        print("| (none) | | |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
