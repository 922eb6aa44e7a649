"""Quick per-k timing of the resident engine (fixed 200 iterations, 148 restarts, C2 shape).
usage: perk_quick.py [k ...]   -> JSON lines"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

ks = [int(a) for a in sys.argv[1:]] or list(range(2, 11))
X = synth.mixture(1000, 200, 5, seed=2015)
n, m = X.shape
with nb.Context(0) as ctx:
    ctx.set_X(X)
    peak = 37.03
    for rep in range(2):
        for k in ks:
            b = ctx.batch(k, 148)
            b.init_random(2015)
            ctx.solve([b], nb.default_params(maxiter=200))
            ms = ctx.last_solve_ms
            b.close()
            if rep:
                tf = 8.0 * n * m * k * 148 * 200 / ms / 1e9
                print(json.dumps(dict(k=k, ms=round(ms, 3), us_per_iter=round(ms * 1e3 / 200, 2), tflops=round(tf, 2),
                                      frac=round(tf / peak, 4), plan=os.environ.get("NMFK_DMMA_PLAN", "auto"))))
