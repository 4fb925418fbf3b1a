#!/usr/bin/env python
"""Development: turn an .ncu-rep into the two files kept under profiles/ per kernel --
<out>_summary.csv (the metrics DESIGN.md / bench.py quote) and <out>_details.txt (ncu --page details).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_<name> [kernel-regex]
"""
import csv
import io
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    extra = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + extra, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out + "_summary.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % k for k in range(len(rows) - 2)])
        for name in ("Kernel Name",) + KEYS:
            if name in hdr:
                i = hdr.index(name)
                w.writerow([name, units[i]] + [r[i] for r in rows[2:]])
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"] + extra, capture_output=True, text=True).stdout
    with open(out + "_details.txt", "w") as f:
        f.write(det)
    print("wrote", out + "_summary.csv", out + "_details.txt")


if __name__ == "__main__":
    main()
