#!/bin/bash
# chain pairs/s of the task-graph schedule for prep:align thread counts (COMBOS="4:3 6:4"); lanes mode for reference
cd "$(dirname "$0")/.."
one() {
  python bench.py --steps 3 --no-cpu-baseline --chain-reps ${REPS:-5} 2>gpurun_out/chain_sweep2.err | \
   python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['extra']['chain']; print('$1', 'pairs/s', round(c['pairs_per_sec'],1), 'ms/chain', round(c['seconds_per_chain']*1e3,2), c['schedule'], [round(x*1e3,1) for x in c['rank0_seconds_per_repetition']], c.get('rank0_device_allocations_per_repetition'), 'iters', c['iterations_total'])" || tail -5 gpurun_out/chain_sweep2.err
}
mkdir -p gpurun_out
if [ -z "$ONLY_DAG" ]; then LC3D_CHAIN_MODE=lanes one "lanes3"; fi
for pa in ${COMBOS:-2:2 3:2 4:2 4:3 6:3 6:4 8:4}; do LC3D_CHAIN_PREP=${pa%:*} LC3D_CHAIN_ALIGN=${pa#*:} one "dag $pa"; done
