"""Source-sharded single-pair ICP (second mode; include/lc3d.h lc3d_shard_*): two
processes exchange their estimator sums through CUDA-IPC-mapped peer memory inside the solve
kernel.  Runs on whatever the box has: two GPUs when present, otherwise both processes share GPU 0
(the IPC mapping and the in-kernel polling are the same; the GPU time-slices the two contexts)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_pair_matches_single_gpu_run():
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env["LC3D_SHARD_ONE_GPU"] = "1"
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "scripts", "shard_check.py"), "2e5"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["parity_ok"] is True, out
    assert out["world"] == 2 and out["mode1"]["iterations"] >= 3 and out["mode0"]["iterations"] >= 3
    assert out["mode1"]["max_abs_T_diff"] < 1e-6 and out["mode0"]["max_abs_T_diff"] < 1e-6
