"""Chain sharding / pose composition (host logic) incl. a world_size-2 gloo run on CPU.
The per-pair aligner is injected; here the CPU oracle plays that role (tests may use it)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from lowcost3dreconstruction_b200 import chain, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_pairs_partition():
    for world in (1, 2, 3, 4, 8):
        seen = []
        for r in range(world):
            seen += chain.shard_pairs(35, world, r)
        assert sorted(seen) == list(range(1, 36))
        sizes = [len(chain.shard_pairs(35, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_compose_and_matrix_files(tmp_path):
    Ts = [synth.rigid(0, 10, 0, [0.01 * i, 0, 0]) for i in range(1, 5)]
    G = chain.compose_chain(Ts)
    assert np.allclose(G[0], np.eye(4)) and np.allclose(G[2], Ts[0] @ Ts[1])
    p = str(tmp_path / "m.txt")
    chain.write_matrix_file(p, G[4])
    assert np.allclose(chain.read_matrix_file(p), G[4], atol=1e-8)
    assert len(open(p).read().split()) == 16  # transform -t reads 16 whitespace-separated numbers


def test_record_roundtrip():
    res = dict(transformation=np.arange(16, dtype=np.float32).reshape(4, 4), fitness=1.5e-5, iterations=9,
               converged=True, state=4)
    back = chain.unpack_record(chain.pack_record(res))
    assert np.array_equal(back["transformation"], res["transformation"]) and back["iterations"] == 9
    assert back["converged"] and back["state"] == 4 and back["fitness"] == 1.5e-5


WORKER = r'''
import os, sys, json
import numpy as np
import pytest
sys.path.insert(0, os.environ["LC3D_ROOT"])
import torch.distributed as dist
from lowcost3dreconstruction_b200 import chain, synth
from oracle import oracle as orc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
views = {}
def get_view(v):
    return synth.kinect_view(v, step_deg=4.0, scale=0.12, backdrop="panel")
def align(s, t):
    return orc.icp_align(s, t, 0.03, 30)
out = chain.register_chain(5, get_view, align, rank, world)
np.save(os.path.join(os.environ["LC3D_OUT"], f"poses_{world}_{rank}.npy"), np.stack(out["pose"]))
dist.destroy_process_group()
'''


def run_world(world, out):
    script = os.path.join(out, "worker.py")
    open(script, "w").write(WORKER)
    env = dict(os.environ, LC3D_ROOT=ROOT, LC3D_OUT=out, OMP_NUM_THREADS="1")
    if world == 1:
        env.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
        subprocess.check_call([sys.executable, script], env=env, timeout=600)
    else:
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                               "--master-addr", "127.0.0.1", "--master-port", "29612", script], env=env, timeout=900)


def test_gloo_world2_matches_world1():
    with tempfile.TemporaryDirectory() as out:
        run_world(1, out)
        run_world(2, out)
        p1 = np.load(os.path.join(out, "poses_1_0.npy"))
        p20 = np.load(os.path.join(out, "poses_2_0.npy"))
        p21 = np.load(os.path.join(out, "poses_2_1.npy"))
        # pairs are independent: identical results for any process count, on every rank
        assert np.array_equal(p1, p20) and np.array_equal(p20, p21)
        # and the chain recovers the turntable: view k is rotated k*4 degrees about +Y
        for k in range(1, 5):
            R = p1[k][:3, :3]
            ang = np.degrees(np.arctan2(R[2, 0], R[0, 0]))
            assert abs(ang - 4.0 * k) < 0.5 * k + 0.3, (k, ang)  # coarse 4.6k-pt views, PCL stops on rel-MSE 1e-3


def test_align_pairs_prefetch_matches_serial():
    """chain.align_pairs: every needed view is fetched exactly once and released exactly once, the
    records are the same with views prepared ahead on the worker thread (prefetch 1, 2) as serially."""
    import threading
    import time

    def run(pairs, prefetch):
        gets, rel, threads = [], [], set()

        def get_view(v):
            time.sleep(0.002)
            gets.append(v)
            threads.add(threading.current_thread().name)
            return ("view", v)

        def align(s, t):
            assert s[1] == t[1] + 1
            time.sleep(0.002)
            return dict(transformation=np.eye(4) * s[1], fitness=0.5 * s[1], iterations=3, converged=True, state=1)

        local = np.zeros((12, chain.RECORD))
        chain.align_pairs(pairs, get_view, align, local, prefetch=prefetch, release=lambda d: rel.append(d[1]))
        return local, gets, rel, threads

    for pairs in ([1, 2, 3, 4, 5], [4, 5, 6], [9], []):
        ref, gets0, rel0, _ = run(pairs, 0)
        need = sorted(set(pairs) | {p - 1 for p in pairs})
        assert sorted(gets0) == need and sorted(rel0) == need
        for pf in (1, 2, 5):
            loc, gets, rel, threads = run(pairs, pf)
            assert np.array_equal(loc, ref)
            assert sorted(gets) == need and sorted(rel) == need
            if pairs:
                assert threads and threading.current_thread().name not in threads  # prepared off-thread


def test_align_pairs_dag_each_view_once_and_same_records():
    """chain.align_pairs_dag: every needed view is prepared exactly once and released exactly once
    (after its last pair), records equal to the serial order, whatever the thread counts."""
    import threading
    import time

    def run(pairs, n_prep, n_align):
        gets, rel, busy, lock = [], [], {"max": 0, "now": 0}, threading.Lock()

        def prep():
            def get_view(v):
                with lock:
                    busy["now"] += 1
                    busy["max"] = max(busy["max"], busy["now"])
                time.sleep(0.003)
                with lock:
                    busy["now"] -= 1
                gets.append(v)
                return ("view", v)
            return get_view

        def align():
            def f(s, t):
                assert s[1] == t[1] + 1 and s[1] not in rel and t[1] not in rel  # never released while in use
                time.sleep(0.002)
                return dict(transformation=np.eye(4) * s[1], fitness=0.5 * s[1], iterations=3, converged=True, state=1)
            return f

        local = np.zeros((12, chain.RECORD))
        chain.align_pairs_dag(pairs, [prep() for _ in range(n_prep)], [align() for _ in range(n_align)], local,
                              release=lambda d: rel.append(d[1]))
        return local, gets, rel, busy["max"]

    for pairs in ([1, 2, 3, 4, 5, 6, 7], [4, 5], [9], []):
        need = sorted(set(pairs) | {p - 1 for p in pairs})
        ref = np.zeros((12, chain.RECORD))
        chain.align_pairs(pairs, lambda v: ("view", v),
                          lambda s, t: dict(transformation=np.eye(4) * s[1], fitness=0.5 * s[1], iterations=3,
                                            converged=True, state=1), ref)
        for n_prep, n_align in ((1, 1), (2, 2), (4, 3)):
            loc, gets, rel, maxbusy = run(pairs, n_prep, n_align)
            assert np.array_equal(loc, ref)
            assert sorted(gets) == need and sorted(rel) == need
            assert maxbusy <= n_prep
            if len(need) >= 4 and n_prep >= 2:
                assert maxbusy >= 2  # views really are prepared concurrently

    def boom(v):
        raise RuntimeError("prepare failed")
    with pytest.raises(RuntimeError):
        chain.align_pairs_dag([1, 2], [boom], [lambda s, t: {}], np.zeros((3, chain.RECORD)))


def test_align_pairs_lanes_split_and_records():
    """chain.align_pairs_lanes: contiguous sub-blocks, one per lane, each lane's callables used only
    by its own thread; records equal to the serial order; one extra view per additional lane."""
    import threading

    def lane():
        seen_threads, gets = set(), []

        def get_view(v):
            seen_threads.add(threading.current_thread().name)
            gets.append(v)
            return ("view", v)

        def align(s, t):
            seen_threads.add(threading.current_thread().name)
            assert s[1] == t[1] + 1
            return dict(transformation=np.eye(4) * s[1], fitness=0.5 * s[1], iterations=3, converged=True, state=1)

        return (get_view, align, lambda d: None), seen_threads, gets

    pairs = list(range(1, 12))
    ref = np.zeros((12, chain.RECORD))
    chain.align_pairs(pairs, lambda v: ("view", v),
                      lambda s, t: dict(transformation=np.eye(4) * s[1], fitness=0.5 * s[1], iterations=3,
                                        converged=True, state=1), ref)
    for n_lanes in (1, 2, 3):
        lanes = [lane() for _ in range(n_lanes)]
        local = np.zeros((12, chain.RECORD))
        chain.align_pairs_lanes(pairs, [ln[0] for ln in lanes], local, prefetch=1)
        assert np.array_equal(local, ref)
        assert sum(len(ln[2]) for ln in lanes) == len(pairs) + n_lanes  # border views prepared once per lane
        blocks = [sorted(ln[2]) for ln in lanes]
        for b in blocks:
            assert b == list(range(b[0], b[-1] + 1))  # contiguous
