"""Short, fixed-iteration cases for ncu captures (one kernel launch each).
usage: ncu_case.py resident K [R] | tiled | tc n m k R"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

mode = sys.argv[1]
with nb.Context(0) as ctx:
    if mode == "resident":
        k = int(sys.argv[2])
        R = int(sys.argv[3]) if len(sys.argv) > 3 else 148
        ctx.set_X(synth.mixture(1000, 200, 5, seed=2015))
        b = ctx.batch(k, R)
        b.init_random(2015)
        ctx.solve([b], nb.default_params(maxiter=30))
        print("resident k=%d R=%d ms=%.3f" % (k, R, ctx.last_solve_ms))
    elif mode == "tc":  # tcgen05 Float32 tiled pass: tc n m k R
        n, m, k, R = (int(v) for v in sys.argv[2:6])
        ctx.set_X(synth.mixture(n, m, 8, seed=3, dtype=np.float32))
        b = ctx.batch(k, R)
        b.init_random(1)
        ctx.solve([b], nb.default_params(maxiter=2, engine=2))
        print("tc tiled ms=%.3f" % ctx.last_solve_ms)
    elif mode == "td":  # DMMA Float64 tiled pass: td n m k R
        n, m, k, R = (int(v) for v in sys.argv[2:6])
        ctx.set_X(synth.mixture(n, m, 8, seed=3))
        b = ctx.batch(k, R)
        b.init_random(1)
        ctx.solve([b], nb.default_params(maxiter=2, engine=2))
        print("dmma tiled ms=%.3f" % ctx.last_solve_ms)
    else:
        ctx.set_X(synth.mixture(20000, 1000, 8, seed=3))
        b = ctx.batch(16, 16)
        b.init_random(1)
        ctx.solve([b], nb.default_params(maxiter=3))
        print("tiled ms=%.3f" % ctx.last_solve_ms)
