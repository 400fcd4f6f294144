#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libsbp_b200.so (cuobjdump -sass): which kernels use the FP64 tensor cores (DMMA), TMA tensor
loads / stores / L2 prefetches (UTMALDG / UTMASTG / UTMAPF), bulk copies (UBLKCP), mbarriers (SYNCS), 256-bit global accesses.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "summationbypartsoperatorsextra.jl_b200", "libsbp_b200.so")
KEYS = ["DMMA", "DFMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "LDG.E.256", "STG.E.256", "LDG.E.128", "STG.E.128",
        "LDS", "STS", "LDL", "STL", "ATOMS", "RED", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for k in KEYS:
                if "." in k:  # LDG.E.256 etc.: base opcode + access width anywhere among the modifiers
                    base, width = k.split(".")[0], k.split(".")[-1]
                    if op.startswith(base + ".") and ("." + width) in op:
                        cur[k] += 1
                elif op == k or op.startswith(k + "."):
                    cur[k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode counts per kernel of {os.path.basename(LIB)} (sm_100a); {len(kernels)} kernels; columns: total " + " ".join(KEYS))
    fam = collections.OrderedDict()
    for (mangled, c), name in zip(kernels.items(), names):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sbp::", "")
        fam.setdefault(re.sub(r"<.*", "", short), []).append((short, c))
    for f, items in fam.items():
        agg = collections.Counter()
        for _, c in items:
            agg.update(c)
        print(f"\n== {f}: {len(items)} instantiation(s); sum: total {agg['total']} " + " ".join(f"{k}={agg[k]}" for k in KEYS if agg[k]))
        for short, c in items[:6]:
            print(f"   {short[:70]:70s} total {c['total']:6d} " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
        if len(items) > 6:
            print(f"   ... {len(items) - 6} more instantiations")


if __name__ == "__main__":
    main()
