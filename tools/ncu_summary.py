#!/usr/bin/env python
"""Summarise an ncu report (--set full) as a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_prof.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]  # fmt: skip


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        print(f"kernel: {d['Kernel Name'][0]}")
        print(f"grid {d['Grid Size'][0]} block {d['Block Size'][0]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:72s} {d[k][0]:>18s} {d[k][1]}")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[h][0]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warps stalled per issue-active cycle (top 7): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]))
        try:
            f = lambda k: float(d[f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed"][0])
            lanes = f("dfma") + f("dmul") + f("dadd")
            flop = 2 * f("dfma") + f("dmul") + f("dadd")
            hz = float(d["sm__cycles_elapsed.avg.per_second"][0]) * {"Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}.get(d["sm__cycles_elapsed.avg.per_second"][1], 1e9)
            print(f"  executed FP64: {lanes:.0f} thread-inst/cycle of 9472 (148 SM x 64 lanes) = {100*lanes/9472:.1f}% of the FP64 lanes; "
                  f"{flop:.0f} flop/cycle = {flop*hz/1e12:.2f} TFLOP/s at {hz/1e9:.2f} GHz (FMA = 2)")
        except (KeyError, ValueError):
            pass
        print()


if __name__ == "__main__":
    main()
