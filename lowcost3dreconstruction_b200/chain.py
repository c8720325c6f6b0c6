"""View-chain registration sharded over the GPUs of one box.

The reference chains the turntable views pairwise (scripts/alignment.sh:106-113: view i is
registered against view i-1 and the transform is applied to all later views, i.e. cumulative
composition; the fine-alignment step at :123-126 is the TODO this fills).  Pairs share no
state, so the pairs are split into contiguous blocks, one per rank, every rank aligns its pairs on
its own GPU, and ONLY the per-pair records (4x4 matrix, fitness, iterations, flags: 20
doubles) are exchanged — one small all-reduce over NCCL/NVLink (gloo in CPU tests).  Rank 0
composes G_0 = I, G_p = G_{p-1} . T_p in float64 and writes `transform -t`-readable matrix
files (pcl_tools/transform.cpp:68-81: 16 whitespace-separated numbers, row-major).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

RECORD = 20  # 16 matrix entries, fitness, iterations, converged, state


def shard_pairs(n_pairs: int, world: int, rank: int) -> list[int]:
    """Pairs are numbered 1..n_pairs (pair p registers view p onto view p-1).  Each rank gets a
    CONTIGUOUS block (sizes differ by at most one): consecutive pairs share a view, so a rank
    loads / preprocesses only len(block)+1 views instead of 2*len(block)."""
    base, extra = divmod(n_pairs, world)
    start = rank * base + min(rank, extra)
    size = base + (1 if rank < extra else 0)
    return list(range(start + 1, start + size + 1))


def compose_chain(pair_transforms: Sequence[np.ndarray]) -> list[np.ndarray]:
    """G_0 = I, G_p = G_{p-1} . T_p (float64): pose of view p in view 0's frame."""
    G = [np.eye(4)]
    for T in pair_transforms:
        G.append(G[-1] @ np.asarray(T, dtype=np.float64))
    return G


def pack_record(res: dict) -> np.ndarray:
    r = np.zeros(RECORD, dtype=np.float64)
    r[:16] = np.asarray(res["transformation"], dtype=np.float64).reshape(16)
    r[16] = res["fitness"]
    r[17] = res["iterations"]
    r[18] = float(res["converged"])
    r[19] = res["state"]
    return r


def unpack_record(r: np.ndarray) -> dict:
    return dict(transformation=r[:16].reshape(4, 4).astype(np.float32), fitness=float(r[16]),
                iterations=int(r[17]), converged=bool(r[18]), state=int(r[19]))


def exchange_records(local: np.ndarray, device=None) -> np.ndarray:
    """local: (n_pairs, RECORD) with this rank's rows filled and zeros elsewhere.  Every pair is
    owned by exactly one rank, so a SUM all-reduce is the gather."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # single process without torch
        return local
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    t = torch.from_numpy(np.ascontiguousarray(local))
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def register_chain(n_views: int, get_view: Callable[[int], object], align: Callable[[object, object], dict],
                   rank: int = 0, world: int = 1, device=None) -> dict:
    """Registers views 1..n_views-1 pairwise (view p onto view p-1).

    get_view(v): returns view v's cloud (called only for the views this rank needs).
    align(source, target): returns the result dict of `api.icp_align` (or a compatible one).
    Returns, on every rank: dict(pair=[records 1..], pose=[G_0..G_{n-1}])."""
    n_pairs = n_views - 1
    local = np.zeros((n_pairs, RECORD), dtype=np.float64)
    cache: dict[int, object] = {}

    def view(v):
        if v not in cache:
            cache[v] = get_view(v)
        return cache[v]

    for p in shard_pairs(n_pairs, world, rank):
        local[p - 1] = pack_record(align(view(p), view(p - 1)))
        cache.pop(p - 1, None)
    allrec = exchange_records(local, device)
    pairs = [unpack_record(r) for r in allrec]
    poses = compose_chain([pr["transformation"] for pr in pairs])
    return dict(pair=pairs, pose=poses)


def write_matrix_file(path: str, T: np.ndarray) -> None:
    """4x4 text file readable by `transform -t` (pcl_tools/transform.cpp:68-81)."""
    T = np.asarray(T, dtype=np.float64).reshape(4, 4)
    with open(path, "w") as f:
        for row in T:
            f.write(" ".join(f"{v:.9g}" for v in row) + "\n")


def read_matrix_file(path: str) -> np.ndarray:
    with open(path) as f:
        vals = [float(x) for x in f.read().split()]
    if len(vals) < 16:
        raise ValueError(f"{path}: expected 16 numbers, found {len(vals)}")
    return np.array(vals[:16], dtype=np.float64).reshape(4, 4)
