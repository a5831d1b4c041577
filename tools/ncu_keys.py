#!/usr/bin/env python
"""Key numbers of one kernel from `ncu -i rep --page raw --csv`: duration, occupancy, pipes, L1 data pipe, DRAM, stalls."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("##", d.get("Kernel Name", "")[:110])
    for k in KEYS:
        if k in d:
            print(f"  {k:82s} {d[k]:>16s} {units[hdr.index(k)]}")
    for h in hdr:
        if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and "ratio" in h and "not_issued" not in h:
            v = float(d[h] or 0)
            if v >= 0.1:
                print(f"  {h[len('smsp__average_warps_issue_stalled_'):].replace('_per_issue_active.ratio',''):82s} {v:16.3f}")
