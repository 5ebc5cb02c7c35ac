"""Summarise ncu raw-page CSV exports: one block per distinct kernel."""
import csv, sys
WANT = ['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread',
 'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps',
 'sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum',
 'lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct',
 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct',
 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_imc_miss_per_warp_active.pct',
 'smsp__cycles_active.avg','sm__cycles_elapsed.max','smsp__thread_inst_executed_per_inst_executed.ratio']
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in data:
        name = r[idx['Kernel Name']]
        key = name[:40]
        if key in seen:
            continue
        seen.add(key)
        print('=' * 10, name[:70])
        for w in WANT:
            if w in idx:
                print(f"  {w:78s} {r[idx[w]]:>16s} {units[idx[w]]}")
