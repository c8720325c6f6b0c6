#!/bin/bash
# ncu launch list of the 36-view chain (one repetition is enough): kernel time by name
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chain_launches.csv python scripts/chain_profile.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/chain_launches.csv")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) > vi:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][-60:]
        agg[name] += v
        cnt[name] += 1
tot = sum(agg.values())
print(f"total kernel time {tot/1e6:.2f} ms over {sum(cnt.values())} launches (3 chain repetitions)")
for k, v in agg.most_common(40):
    print(f"{v/1e3/3:9.1f} us/chain {cnt[k]//3:5d} launches/chain {v/cnt[k]/1e3:8.2f} us each  {k}")
PY
