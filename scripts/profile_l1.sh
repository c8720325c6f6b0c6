#!/bin/bash
# One ncu --set full capture of the ICP iteration kernels of one alignment; dumps every raw metric of an
# early (launch 2) and a converged (launch 7) iteration as "name = value unit" lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LIB=${1:-}
[ -n "$LIB" ] && export LC3D_LIB=$(realpath $LIB)
ncu --set full --clock-control none --import-source on -k regex:icp_iteration -c 12 \
  -o gpurun_out/l1_icp -f python scripts/dev_profile_icp.py 1 > gpurun_out/l1_ncu.log 2>&1
tail -2 gpurun_out/l1_ncu.log
ncu -i gpurun_out/l1_icp.ncu-rep --page raw --csv > gpurun_out/l1_icp_raw.csv
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/l1_icp_raw.csv")))
hdr, units = rows[0], rows[1]
want = ("l1tex", "lsu", "lts__t_sector_hit", "smsp__warp_issue_stalled", "sm__inst_executed_pipe", "gpu__time",
        "sm__cycles_active", "sm__throughput", "smsp__inst_executed.sum", "sm__warps_active", "smsp__issue_active",
        "idc__", "smsp__average_warp", "smsp__thread_inst", "shared", "local")
for li in (1, 6):
    r = rows[2 + li]
    with open(f"gpurun_out/l1_icp_launch{li}.txt", "w") as f:
        for h, u, v in zip(hdr, units, r):
            if any(w in h for w in want):
                f.write(f"{h} = {v} {u}\n")
PY
rm -f gpurun_out/l1_icp_raw.csv
ls -la gpurun_out/
