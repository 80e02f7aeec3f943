#!/usr/bin/env python
"""Text summary of an .ncu-rep capture (ncu --set full): per kernel launch the duration, launch shape, occupancy, issue
rate, DRAM bytes, shared-memory wavefronts / bank conflicts, executed instructions and the warp-stall breakdown.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        get = dict(zip(hdr, r))
        print(f"== {get.get('Kernel Name', '?')}   (launch id {get.get('ID', '?')})")
        for k in WANT:
            if k in get and get[k] != "":
                print(f"   {k:78s} {get[k]:>18s} {units[hdr.index(k)]}")
        st = sorted(((float(get[h] or 0), h) for h in stall), reverse=True)[:6]
        print("   top warp stalls (warps stalled per issue-active cycle): " +
              ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in st))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
