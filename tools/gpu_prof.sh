#!/bin/bash
# Phase trace + quick counters of one config (default C3): rebuilds the library with BK_PROFILE_PHASES on the box,
# runs tools/phases.py, rebuilds the normal library, then a few ncu counters of one launch.
cfg=${1:-C3}; k=${2:-1}; R=${3:-32}
mkdir -p gpurun_out
BK_PROFILE_PHASES=1 python -m barkit_b200.build --force >/dev/null 2>&1
echo "== phases $cfg k=$k"; timeout -k 5 200 python tools/phases.py 4000000 $R $cfg $k 2>&1 | tail -24
python -m barkit_b200.build --force >/dev/null 2>&1
echo "== speed"; BK_ONLY=$cfg timeout -k 5 200 python tools/config_speed.py 4000000 2>&1 | tail -3
echo "== ncu"; BK_ONLY=$cfg timeout -k 5 300 ncu --clock-control none -k regex:extract_kernel -s 2 -c 1 --metrics sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active python tools/config_speed.py 4000000 2>&1 | grep -E "icc|gcc|inst_executed|time_duration|wavefronts|dram__|lts__|stalled|launch__|warps_active"
