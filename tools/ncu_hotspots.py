#!/usr/bin/env python
"""Top stall-sample instructions of an .ncu-rep captured with --import-source on (SASS view).

    python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            data.append((int(r[ci["# Samples"]]), r))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"total samples {tot}")
    agg = {s: 0 for s in stalls}
    for n, r in data:
        for s in stalls:
            try:
                agg[s] += int(r[ci[s]])
            except ValueError:
                pass
    print("stall mix:", ", ".join(f"{s[6:]}={100 * v / tot:.1f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    order = {id(r): k for k, (_, r) in enumerate(data)}
    for n, r in sorted(data, key=lambda x: -x[0])[:top]:
        main_stall = max(stalls, key=lambda s: int(r[ci[s]] or 0))
        print(f"{100 * n / tot:5.1f}%  #{order[id(r)]:4d}  {main_stall[6:]:12s} {r[ci['Source']][:100]}")


if __name__ == "__main__":
    main()
