#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15
echo "== v3 stats mode 1"; LC3D_STATS=1 python scripts/dev_profile_icp.py 1 2 2>&1 | grep "lc3d stats" | tail -22
run() { echo "== $*"; env "$@" python scripts/dev_profile_icp.py 1 6 2>&1 | tail -1 | cut -c1-200; }
run LC3D_ICP_V1=1
run LC3D_X=0
run LC3D_LISTS=0
run LC3D_LIB=$PWD/lowcost3dreconstruction_b200/csrc/liblc3d_mb3.so
for cf in 1.5 2 2.5 3 4; do for xs in 1 2 8; do run LC3D_CELL_FACTOR=$cf LC3D_XSUB=$xs; done; done
run LC3D_MU_KAPPA=1.5
run LC3D_MU_KAPPA=3
run LC3D_MU_MAX=2.5
run LC3D_MU_MAX=1.0
echo "== mode 0"; python scripts/dev_profile_icp.py 0 6 2>&1 | tail -1 | cut -c1-200
