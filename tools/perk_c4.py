"""Per-k timing of the tiled engine at the C4 shape (100000 x 2000 Float64, 32 restarts, 3 iterations): where the k = 2:32 sweep
spends its time.  usage: perk_c4.py [k ...]  -> JSON lines"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

ks = [int(a) for a in sys.argv[1:]] or list(range(2, 33))
n, m, R, iters = 100000, 2000, 32, 3
X = synth.mixture(n, m, 8, seed=2015, dtype=np.float64)
peak = 37.0
tot = 0.0
with nb.Context(0) as ctx:
    ctx.set_X(X)
    for rep in range(2):
        for k in ks:
            b = ctx.batch(k, R)
            b.init_random(2015)
            ctx.solve([b], nb.default_params(maxiter=iters, engine=2))
            ms = ctx.last_solve_ms
            b.close()
            if rep:
                tf = 8.0 * n * m * k * R * iters / ms / 1e9
                tot += ms
                print(json.dumps(dict(k=k, ms=round(ms, 2), rit_per_s=round(R * iters / ms * 1e3, 1), tflops=round(tf, 2), frac=round(tf / peak, 3))))
print(json.dumps(dict(total_ms=round(tot, 1), sweep_rit_per_s=round(len(ks) * R * iters / tot * 1e3, 1))))
