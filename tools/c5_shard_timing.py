"""One GPU's share of the row-sharded C5 run (n_local rows of the 2,000,000 x 1000 Float32 matrix, k = 24, nNMF = 32) through the
sharded code path with a single rank (same kernels as the N-GPU run, no NCCL): per-phase device time of an iteration with
NMFK_TILED_TIMING=1.   usage: c5_shard_timing.py [n_local] [iters]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402

n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m, k, R, k0 = 1000, 24, 32, 24
H0 = np.random.Generator(np.random.Philox(key=2015)).random((k0, m))
W0 = np.random.Generator(np.random.Philox(key=2016)).random((n_local, k0))
X = np.asfortranarray((W0 @ H0).astype(np.float32))
del W0
for sharded in (True, False):
    with nb.Context(0) as ctx:
        if sharded:
            ctx.comm_init(1, 0, None, 0, n_local)
        ctx.set_X(X)
        for it in (3, iters, iters):
            b = ctx.batch(k, R)
            b.init_random(2015)
            ctx.solve([b], nb.default_params(maxiter=it, engine=2))
            ms = ctx.last_solve_ms
            tot = int(b.get(factors=False)["iters"].sum())
            b.close()
        print(json.dumps(dict(n_local=n_local, sharded_path=sharded, iters=iters, ms=ms, value=tot / ms * 1e3)), flush=True)
