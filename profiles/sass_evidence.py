#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libouzelum_b200.so (cuobjdump -sass): python profiles/sass_evidence.py <out.md>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ouzelum_b200", "libouzelum_b200.so")
COLS = [("UBLKCP", ("UBLKCP",)), ("UTMALDG+UTMASTG", ("UTMALDG", "UTMASTG")), ("SYNCS", ("SYNCS",)), ("LDG", ("LDG",)), ("STG", ("STG",)), ("LDS", ("LDS",)), ("STS", ("STS",)),
        ("MUFU", ("MUFU",)), ("FFMA", ("FFMA",)), ("FMUL+FADD", ("FMUL", "FADD")), ("DFMA+DMUL+DADD", ("DFMA", "DMUL", "DADD")),
        ("RED/ATOM", ("RED", "REDG", "ATOMG", "ATOM"))]


def main(out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    with open(out, "w") as f:
        f.write("# SASS evidence (cuobjdump -sass ouzelum_b200/libouzelum_b200.so, sm_100a)\n\n"
                "Blackwell/Hopper async-copy and barrier mnemonics per kernel (TMA 1-D bulk copy = `UBLKCP`, 2-D tensor-map copy = `UTMALDG` / `UTMASTG`, mbarrier = `SYNCS`; no `HMMA`/`UTC*MMA`: "
                "nothing on this path is a dense contraction).\n\n")
        f.write("| kernel | SASS instructions | " + " | ".join(c for c, _ in COLS) + " |\n|---|---:|" + "---:|" * len(COLS) + "\n")
        for name, c in kernels.items():
            short = re.sub(r"^_ZN3ozl\d+", "", name)[:44]
            f.write(f"| `{short}` | {c['_total']} | " + " | ".join(str(sum(c[k] for k in ks)) for _, ks in COLS) + " |\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_evidence.md"))
