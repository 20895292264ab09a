#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/.
usage: scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r1_k2.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.per_cycle_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def traffic(argv):
    """--traffic kernel_regex batch rep [rep...]: dram bytes (read + write) per launch, summed over the
    launches of one scan (all captured launches of that kernel are taken as one scan's phases)."""
    import json
    import re
    out = {}
    for spec in argv:
        name, batch, rep, how, *key = spec.split(":")  # how = sum (launches are the phases of one scan) | mean;
        key = key[0] if key else name                  # optional 5th field: the key of the entry
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, data = rows[0], rows[1], rows[2:]
        col = {k: i for i, k in enumerate(head)}
        tot, n, us = 0.0, 0, 0.0
        for r in data:
            if not re.search(name, r[col["Kernel Name"]]):
                continue
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = float(r[col[k]]), units[col[k]].lower()
                tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            n += 1
        out[key] = {"batch": int(batch), "dram_bytes_per_launch": tot if how == "sum" else tot / max(n, 1),
                     "captured_kernel_launches": n,
                     "source": rep + " (ncu --set full --clock-control none; " +
                               ("the captured launches are the bootstrap + phases of one scan, summed)" if how == "sum"
                                else "mean over the captured launches)")}
    print(json.dumps(out, indent=1))


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2:])
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(head)}
    print(f"# {rep}: ncu --set full --clock-control none, {len(data)} launch(es)")
    for r in data:
        print(f"\n## {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k in WANT:
            if k in col:
                print(f"{k:95s} {r[col[k]]:>18s} {units[col[k]]}")


if __name__ == "__main__":
    main()
