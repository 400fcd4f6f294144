#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_config2_ncu.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    lines = [f"# condensed from {rep} (ncu --set full --clock-control none); one block per captured launch"]
    for r in rows[2:]:
        lines.append(f"\n== {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:85s} {r[i]:>18s} {units[i]}")
        stalls = [(float(r[i]), h[len(STALLS):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
                  if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and r[i]]
        stalls.sort(reverse=True)
        lines.append("top stall reasons (warps stalled per issue-active cycle): " +
                     ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
