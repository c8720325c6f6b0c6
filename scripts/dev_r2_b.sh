#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== v2 per-iteration"; LC3D_STATS=1 python scripts/dev_profile_icp.py 1 2 2>&1 | grep "kernel" | tail -10
echo "== v1 per-iteration"; LC3D_ICP_V1=1 LC3D_STATS=1 python scripts/dev_profile_icp.py 1 2 2>&1 | grep "kernel" | tail -10
timeout 600 ncu --set full --import-source on --clock-control none -k regex:icp_iter2 -c 9 -f -o gpurun_out/r02a_iter2 python scripts/dev_profile_icp.py 1 1 > gpurun_out/ncu_a.log 2>&1
ncu -i gpurun_out/r02a_iter2.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,sm__cycles_active.avg 2>&1 | tail -12
