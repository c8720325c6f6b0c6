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


def align_pairs(pairs: Sequence[int], get_view: Callable[[int], object], align: Callable[[object, object], dict],
                local: np.ndarray, prefetch: int = 0, release: Callable[[object], None] | None = None) -> None:
    """Aligns `pairs` (a contiguous block, ascending) and writes their records into `local`.

    prefetch = 0: everything on the calling thread, view v fetched right before its first pair.
    prefetch > 0: get_view runs on ONE worker thread up to `prefetch` views ahead of the pair that is
    being aligned — with get_view bound to its own lc3d context (its own stream and scratch), the
    per-view passes (VoxelGrid, SOR, normals: short kernels separated by host round trips) overlap
    the ICP loop of the previous pair on the same GPU (the scripts/alignment.sh:99-113 stages as a
    two-stage software pipeline).  release(view), if given, runs on the same worker (a view's
    buffers go back to the pool of the context that made them, which is not thread-safe)."""
    pairs = list(pairs)
    if not pairs:
        return
    views = sorted({p for p in pairs} | {p - 1 for p in pairs})
    if prefetch <= 0:
        cache: dict[int, object] = {}
        for p in pairs:
            for v in (p, p - 1):
                if v not in cache:
                    cache[v] = get_view(v)
            local[p - 1] = pack_record(align(cache[p], cache[p - 1]))
            old = cache.pop(p - 1)
            if release:
                release(old)
        if release:
            for v in cache.values():
                release(v)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=1) as pool:
        fut, nxt = {}, 0

        def top_up(upto):  # keep views[..upto] submitted
            nonlocal nxt
            while nxt < len(views) and nxt <= upto:
                fut[views[nxt]] = pool.submit(get_view, views[nxt])
                nxt += 1

        top_up(1 + prefetch)
        for i, p in enumerate(pairs):
            tgt, src = fut[p - 1].result(), fut[p].result()
            top_up(i + 2 + prefetch)  # views[i + 1] = p is in use; stay `prefetch` views ahead
            local[p - 1] = pack_record(align(src, tgt))
            del fut[p - 1]
            if release:
                pool.submit(release, tgt)
        if release:
            for f in fut.values():
                pool.submit(release, f.result())


def align_pairs_lanes(pairs: Sequence[int], lanes: Sequence[tuple], local: np.ndarray, prefetch: int = 0) -> None:
    """Several host threads on ONE GPU: `pairs` is split into len(lanes) contiguous sub-blocks, lane i =
    (get_view, align, release) bound to its own lc3d contexts runs align_pairs on sub-block i in its
    own thread.  A turntable view after VoxelGrid + SOR is ~50k points: one pair keeps the 148 SMs
    busy for a fraction of each launch, so independent pairs on independent streams fill the rest.
    Costs len(lanes) - 1 extra view preparations (the sub-blocks' border views)."""
    pairs = list(pairs)
    lanes = list(lanes)[:max(1, len(pairs))]
    if len(lanes) <= 1:
        gv, al, rl = lanes[0]
        return align_pairs(pairs, gv, al, local, prefetch, rl)
    import threading
    errors: list[BaseException] = []

    def run(i):
        try:
            sub = shard_pairs(len(pairs), len(lanes), i)  # 1-based positions inside `pairs`
            gv, al, rl = lanes[i]
            align_pairs([pairs[j - 1] for j in sub], gv, al, local, prefetch, rl)
        except BaseException as e:  # noqa: BLE001 - re-raised on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=run, args=(i,), name=f"lc3d-lane-{i}") for i in range(len(lanes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]


def align_pairs_dag(pairs: Sequence[int], prep_fns: Sequence[Callable[[int], object]],
                    align_fns: Sequence[Callable[[object, object], dict]], local: np.ndarray,
                    release: Callable[[object], None] | None = None) -> None:
    """The rank's block as a task graph on ONE GPU: prepare(v) for every needed view (each exactly
    once) on a pool of len(prep_fns) host threads, align(p) as soon as views p and p-1 are ready on a
    pool of len(align_fns) threads.  Every callable is bound to its own lc3d context (own stream, own
    scratch) and is used by one thread at a time.  Unlike the sub-block pipelines of
    align_pairs_lanes no view is prepared twice, and a block of 4-5 pairs (8 GPUs) still keeps several
    streams busy instead of one pipeline that never reaches its steady state.  A view is released
    (release(view), from whichever thread finishes its last pair) once no pair needs it any more."""
    pairs = list(pairs)
    if not pairs:
        return
    import queue
    import threading
    from concurrent.futures import ThreadPoolExecutor
    views = sorted({p for p in pairs} | {p - 1 for p in pairs})
    users = {v: sum(1 for p in pairs if v in (p, p - 1)) for v in views}
    lock = threading.Lock()
    prep_q: "queue.SimpleQueue" = queue.SimpleQueue()
    align_q: "queue.SimpleQueue" = queue.SimpleQueue()
    for f in prep_fns:
        prep_q.put(f)
    for f in align_fns:
        align_q.put(f)

    def do_prep(v):
        f = prep_q.get()
        try:
            return f(v)
        finally:
            prep_q.put(f)

    def do_align(p, fsrc, ftgt):
        src, tgt = fsrc.result(), ftgt.result()
        f = align_q.get()
        try:
            rec = pack_record(f(src, tgt))
        finally:
            align_q.put(f)
        local[p - 1] = rec
        if release:
            for v, obj in ((p, src), (p - 1, tgt)):
                with lock:
                    users[v] -= 1
                    last = users[v] == 0
                if last:
                    release(obj)

    with ThreadPoolExecutor(max_workers=len(prep_fns), thread_name_prefix="lc3d-prep") as pp, \
            ThreadPoolExecutor(max_workers=len(align_fns), thread_name_prefix="lc3d-align") as ap:
        vf = {v: pp.submit(do_prep, v) for v in views}  # ascending: the pairs become ready in order
        af = [ap.submit(do_align, p, vf[p], vf[p - 1]) for p in pairs]
        err = None
        for f in af:
            try:
                f.result()
            except BaseException as e:  # noqa: BLE001 - first error re-raised after the pools drain
                err = err or e
        if err:
            raise err


class NativeChain:
    """The same task graph run by the library's own host threads (include/lc3d.h: lc3d_chain_*): one C
    call per pair block, no interpreter between the tasks.  views: host xyz arrays of consecutive
    views; pair i registers view i+1 onto view i."""

    def __init__(self, device: int = 0, prepare_threads: int = 4, align_threads: int = 3):
        import ctypes as C

        from . import _capi
        self._lib = _capi.load()
        h = C.c_void_p()
        rc = self._lib.lc3d_chain_create(int(device), int(prepare_threads), int(align_threads), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"lc3d_chain_create failed ({rc}): {self._lib.lc3d_last_error(None).decode()}")
        self._h = h

    def run(self, views, leaf_size=0.0, sor_mean_k=0, sor_stddev_mul=1.0, normals_k=0, viewpoint=(0.0, 0.0, 0.0),
            max_correspondence_distance=0.1, max_iterations=50, transformation_epsilon=1e-9,
            euclidean_fitness_epsilon=1e-3, mode=0, warm=False):
        """Returns (list of api.icp_align-style result dicts, one per pair; points per prepared view)."""
        import ctypes as C

        from . import _capi, api
        hcs = [v if isinstance(v, _capi.HostCloud) else _capi.HostCloud(v) for v in views]
        arr = (_capi.Cloud * len(hcs))(*[h.struct for h in hcs])
        pp = _capi.PrepareParams(float(leaf_size), int(sor_mean_k), float(sor_stddev_mul), int(normals_k),
                                 (C.c_float * 3)(*[float(x) for x in viewpoint]))
        ip = _capi.IcpParams(float(max_correspondence_distance), float(transformation_epsilon),
                             float(euclidean_fitness_epsilon), int(max_iterations), int(mode), 1, -1)
        res = (_capi.IcpResult * (len(hcs) - 1))()
        npts = (C.c_int64 * len(hcs))()
        rc = self._lib.lc3d_chain_run(self._h, arr, len(hcs), C.byref(pp), C.byref(ip), res, npts, int(bool(warm)))
        if rc != 0:
            raise RuntimeError(f"lc3d_chain_run failed ({rc}): {self._lib.lc3d_chain_last_error(self._h).decode()}")
        return [api._result_dict(r) for r in res], [int(x) for x in npts]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lc3d_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def register_chain(n_views: int, get_view: Callable[[int], object], align: Callable[[object, object], dict],
                   rank: int = 0, world: int = 1, device=None, prefetch: int = 0,
                   release: Callable[[object], None] | None = None) -> dict:
    """Registers views 1..n_views-1 pairwise (view p onto view p-1).

    get_view(v): returns view v's cloud (called only for the views this rank needs).
    align(source, target): returns the result dict of `api.icp_align` (or a compatible one).
    prefetch / release: see align_pairs (views prepared ahead on a worker thread).
    Returns, on every rank: dict(pair=[records 1..], pose=[G_0..G_{n-1}])."""
    n_pairs = n_views - 1
    local = np.zeros((n_pairs, RECORD), dtype=np.float64)
    align_pairs(shard_pairs(n_pairs, world, rank), get_view, align, local, prefetch, release)
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


# --------------------------------------------------------------------------------------------
# Second multi-GPU mode: ONE large pair, the source sharded over the ranks.

class ShardedPair:
    """Source-sharded ICP of one pair over `world` processes (one GPU each; torch.distributed is the
    control plane: it carries the 64-byte IPC handles once and the barrier between alignments — the
    per-iteration exchange of the estimator sums happens inside the solve kernel over peer memory,
    see include/lc3d.h).  Every rank passes its slice of the source and the WHOLE target."""

    def __init__(self, ctx, max_shard_points: int):
        import ctypes as C

        import torch.distributed as dist
        self.ctx, self.dist = ctx, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        handle = (C.c_ubyte * 64)()
        ctx._check(ctx._lib.lc3d_shard_export(ctx._h, int(max_shard_points), handle), "lc3d_shard_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        blob = b"".join(handles)
        ctx._check(ctx._lib.lc3d_shard_connect(ctx._h, self.rank, self.world, blob), "lc3d_shard_connect")
        dist.barrier()

    def shard(self, n: int) -> slice:
        """this rank's contiguous slice of an n-point source"""
        base, extra = divmod(n, self.world)
        start = self.rank * base + min(self.rank, extra)
        return slice(start, start + base + (1 if self.rank < extra else 0))

    def align(self, src_shard, tgt, max_correspondence_distance=0.1, max_iterations=50, transformation_epsilon=1e-9,
              euclidean_fitness_epsilon=1e-3, mode=0, dump_iteration=-1) -> dict:
        """src_shard / tgt: api.DeviceCloud.  Collective.  Returns the api.icp_align result dict with the
        fitness of the WHOLE pair (shard sums combined with one all_gather_object)."""
        import ctypes as C

        from . import api
        from ._capi import IcpOutputs, IcpParams, IcpResult
        ctx = self.ctx
        p = IcpParams(float(max_correspondence_distance), float(transformation_epsilon), float(euclidean_fitness_epsilon),
                      int(max_iterations), int(mode), 1, int(dump_iteration))
        r, o, out = IcpResult(), IcpOutputs(), {}
        if dump_iteration >= 0:
            out["corr_index"] = np.empty(src_shard.n, dtype=np.int32)
            out["corr_dist2"] = np.empty(src_shard.n, dtype=np.float32)
            o.corr_index, o.corr_dist2 = out["corr_index"].ctypes.data, out["corr_dist2"].ctypes.data
        parts = (C.c_double * 2)()
        self.dist.barrier()  # nobody may still be reading the previous alignment's sums
        ctx._check(ctx._lib.lc3d_icp_align_sharded(ctx._h, src_shard._h, tgt._h, C.byref(p), C.byref(r), C.byref(o), parts),
                   "lc3d_icp_align_sharded")
        out.update(api._result_dict(r))
        allparts = [None] * self.world
        self.dist.all_gather_object(allparts, (parts[0], parts[1]))
        tot, cnt = sum(a for a, _ in allparts), sum(b for _, b in allparts)
        out["fitness"] = tot / cnt if cnt > 0 else float(np.finfo(np.float64).max)
        return out

    def close(self):
        self.ctx._lib.lc3d_shard_close(self.ctx._h)
