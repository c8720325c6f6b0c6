#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "icp" 2>&1 | tail -3
python bench.py > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; tail -2 gpurun_out/g2_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/g2_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print("extra", {k: v for k, v in d["extra"].items() if not isinstance(v, dict)})
print("aos", d["extra"]["e2e_pcl_aos_pageable"]["value"])
print("chain", d["extra"]["chain"]["pairs_per_sec"], d["extra"]["chain"]["seconds_per_chain"], d["extra"]["chain"].get("lanes_per_gpu"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cfg1_point_to_point"]["gpu_iters_per_sec_resident"])
PY
LC3D_NO_PACK=1 python bench.py --steps 10 --no-cpu-baseline --no-chain 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no-pack aos', d['extra']['e2e_pcl_aos_pageable']['value'], 'e2e', d['e2e']['value'])"
