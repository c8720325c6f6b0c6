#!/bin/bash
# chain pairs/s for combinations of lanes x prefetch (one GPU); every setting twice (run-to-run spread)
cd "$(dirname "$0")/.."
for l in ${LANES:-1 2 3 4}; do for pf in ${PREFETCH:-0 1}; do for rep in 1 2; do
  LC3D_CHAIN_LANES=$l LC3D_CHAIN_PREFETCH=$pf python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | \
   python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['extra']['chain']; print('lanes $l prefetch $pf pairs/s', round(c['pairs_per_sec'],1), 'ms/chain', round(c['seconds_per_chain']*1e3,2), 'lanes used', c.get('lanes_per_gpu'), 'iters', c['iterations_total'], 'maxerr', round(c['max_rot_err_deg_vs_truth'],4))"
done; done; done
