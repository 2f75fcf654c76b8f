"""Condense an `ncu --set full` report of tools/ncu_targets.py into the per-kernel metric table kept under profiles/.
    ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/<name>_summary.csv
Every target is launched three times (warm-up + 2); the LAST launch of each kernel variant is reported."""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
LABELS = ["fwd_L1_kernel", "fwd_staged_200KB", "bwd_L1_kernel_mode1", "bwd_staged_tensor_core_mode2",
          "bwd_staged_coarse_REDs_dropped_mode4", "bwd_small_cta_tensor_core_level3_mode5"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
picked = [data[i] for i in range(2, len(data), 3)]          # third launch of each target
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + LABELS[:len(picked)])
w.writerow(["kernel", ""] + [r[ki].split("(")[0].replace("void unnamed>::", "") for r in picked])
for m in METRICS:
    if m in hdr:
        j = hdr.index(m)
        w.writerow([m, units[j]] + [r[j] for r in picked])
