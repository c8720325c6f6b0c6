#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:icp_iter3 -c 9 -f -o gpurun_out/r02b_iter3 python scripts/dev_profile_icp.py 1 1 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
