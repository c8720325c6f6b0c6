#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { echo "== $*"; env "$@" python scripts/dev_profile_icp.py ${MODE:-1} 8 2>&1 | tail -1 | cut -c1-170; }
run LC3D_X=0
run LC3D_RADIX_INDEX=1
run LC3D_X=0
python scripts/dev_chain_probe.py 2>&1 | tail -10
