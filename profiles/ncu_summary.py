#!/usr/bin/env python
"""Key metrics of one `ncu --set full` report as text.  usage: python profiles/ncu_summary.py gpurun_out/prof_X.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"# {path}\nkernel: {name[:150]}")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                print(f"  {h:80s} {v:>16s} {u}")
        rd = float(vals[hdr.index("dram__bytes_read.sum")])
        wr = float(vals[hdr.index("dram__bytes_write.sum")])
        print(f"  traffic (dram read + write)                                                      {rd + wr:16.3f} {units[hdr.index('dram__bytes_read.sum')]}")


if __name__ == "__main__":
    main(sys.argv[1])
