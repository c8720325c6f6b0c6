#!/bin/bash
# several fresh processes (some with LC3D_CHUNK set), distinct outcomes of repeated alignments in each
cd "$(dirname "$0")/.."
nvidia-smi -L; hostname
for e in "" "LC3D_CHUNK=3" "LC3D_CHUNK=2" "" "LC3D_XSUB=8 LC3D_CELL_FACTOR=4" "LC3D_CHUNK=1"; do
  echo "== env [$e]"; env $e python scripts/determinism_check.py 60 2>&1 | grep -v "^$" | tail -7
done
