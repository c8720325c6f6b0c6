#!/bin/bash
# bench at N GPUs (torchrun), JSON line kept under gpurun_out/ ; extra args = env assignments for a second run
cd "$(dirname "$0")/.."
N=$1; shift
mkdir -p gpurun_out
run() {
  tag=$1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 2> gpurun_out/n${N}_bench$tag.err | grep '^{' > gpurun_out/n${N}_bench$tag.json
  python - <<PY
import json
d = json.load(open("gpurun_out/n${N}_bench$tag.json")); c = d["extra"]["chain"]
print("N=${N}$tag value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["e2e"].get("pcie_pinned_gbs"), "chain pairs/s", round(c["pairs_per_sec"], 1),
      "ms/chain", round(c["seconds_per_chain"] * 1e3, 2), "lanes", c["lanes_per_gpu"], c["rank0_seconds_per_repetition"], "identical", d["extra"].get("all_ranks_bit_identical_results"), "by rank", d["extra"].get("ms_per_step_by_rank"))
PY
}
nproc
run ""
if [ -n "$1" ]; then export "$@"; run "_alt"; fi
