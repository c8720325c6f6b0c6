"""Summarise an LC3D_BLOCK_LOG file: per iteration the kernel span, the busy time per SM, the tail."""
import sys
import numpy as np
raw = open(sys.argv[1], "rb").read()
iters, nblk = np.frombuffer(raw[:8], np.int32)
a = np.frombuffer(raw[8:], np.uint64).reshape(iters, nblk, 3).astype(np.int64)
for it in range(iters):
    t0, t1, sm = a[it, :, 0], a[it, :, 1], a[it, :, 2]
    if t0.max() == 0:
        break
    base = t0.min()
    s, e = (t0 - base) / 1e3, (t1 - base) / 1e3
    dur = e - s
    span = e.max()
    # per-SM last finish
    last = np.array([e[sm == k].max() for k in np.unique(sm)])
    # number of blocks still running over time
    order = np.argsort(e)
    print(f"it {it}: span {span:6.1f} us | block dur mean {dur.mean():5.1f} p50 {np.median(dur):5.1f} p90 {np.percentile(dur,90):5.1f} "
          f"p99 {np.percentile(dur,99):5.1f} max {dur.max():5.1f} | last start {s.max():5.1f} | SM finish p10 {np.percentile(last,10):5.1f} "
          f"p50 {np.median(last):5.1f} p90 {np.percentile(last,90):5.1f} | blocks ending in last 5us {int((e > span - 5).sum())} "
          f"| heaviest 5 blocks idx {list(np.argsort(-dur)[:5])} start {[round(float(x),1) for x in s[np.argsort(-dur)[:5]]]}")
