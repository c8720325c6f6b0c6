#!/bin/bash
# Round-2 A/B helper (developer script, run under gpurun): parity first, then loop timings of
# the first- and second-generation iteration kernels on the bench pair.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15 > gpurun_out/ab_pytest.log
cat gpurun_out/ab_pytest.log
for mode in 1 0; do
  echo "== v1 mode $mode"; LC3D_ICP_V1=1 python scripts/dev_profile_icp.py $mode 6 2>&1 | tail -2
  echo "== v2 mode $mode"; python scripts/dev_profile_icp.py $mode 6 2>&1 | tail -2
done
echo "== v2 stats mode 1"; LC3D_STATS=1 python scripts/dev_profile_icp.py 1 2 2>&1 | grep -v "^\[lc3d stats\] it .. kernel" | tail -40
for extra in "$@"; do
  echo "== v2 $extra"; env $extra python scripts/dev_profile_icp.py 1 6 2>&1 | tail -1
done
