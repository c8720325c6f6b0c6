import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from lowcost3dreconstruction_b200 import api
src, tgt = bench.load_pair(0)
ctx = api.Context(0)
for name, fn in (("normals k=30", lambda: api.normals(tgt, 30, ctx=ctx)), ("sor k=50", lambda: api.sor(tgt, 50, 1.0, ctx=ctx)), ("knn k=8", lambda: api.knn(tgt, 8, ctx=ctx))):
    fn(); fn()
    t = time.perf_counter()
    for _ in range(5): fn()
    print(f"{name}: {(time.perf_counter()-t)/5*1e3:.2f} ms (host call)", ctx.grid_info()["cell"])
