#!/bin/bash
# Final round-2 ncu evidence (run under gpurun, one GPU): launch list of the bench command, --set full
# metrics of every icp_* launch of one alignment (+ by-source-line of a converged iteration), the
# k-NN kernels of one chain, and the DRAM traffic figure bench.py reads.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02c_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-chain > gpurun_out/r02c_launches_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:icp_ -c 30 -f -o gpurun_out/r02c_icp \
  python scripts/dev_profile_icp.py 1 1 > gpurun_out/r02c_icp_ncu.log 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
ncu -i gpurun_out/r02c_icp.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02c_icp_kernels_ncu_full.csv 2>&1
python scripts/ncu_by_line.py gpurun_out/r02c_icp.ncu-rep icp_iteration_kernel 7 30 > gpurun_out/r02c_icp_iteration_converged_by_source_line.txt 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:knn_kernel -c 4 --csv --log-file gpurun_out/r02c_knn_kernels.csv \
  python scripts/chain_profile.py > /dev/null 2>&1
rm -f gpurun_out/r02c_icp.ncu-rep
tail -3 gpurun_out/r02c_icp_ncu.log; ls -la gpurun_out | grep r02c
