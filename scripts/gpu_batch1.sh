#!/bin/bash
# one GPU visit: full parity suite, chain sweep, A/B of the cooperative ball walk, e2e breakdown
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g1_tests.log 2>&1; tail -6 gpurun_out/g1_tests.log
echo "== chain sweep"; scripts/chain_sweep.sh 2>&1 | tee gpurun_out/g1_chain_sweep.log
echo "== prep timing (hint / no hint)"
LC3D_PREP_TIMING=1 LC3D_CHAIN_LANES=1 python scripts/chain_profile.py 2>&1 | tail -4
LC3D_NO_GRID_HINT=1 LC3D_PREP_TIMING=1 python scripts/chain_profile.py 2>&1 | tail -4
V=lowcost3dreconstruction_b200/csrc/variants
echo "== ball rows"; python scripts/ab_icp.py --both $V/lib_ball1.so $V/lib_ball2.so default $V/lib_ball4.so 2>&1 | tee gpurun_out/g1_ab_ball.log
echo "== e2e"; python scripts/dev_e2e.py 2>&1 | tee gpurun_out/g1_e2e.log
LC3D_DEFER_NORMALS=0 python scripts/dev_e2e.py 2>&1 | head -2
LC3D_NO_PACK=1 python scripts/dev_e2e.py 2>&1 | sed -n 3,4p
