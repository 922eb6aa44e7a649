"""Variant FRO on the C3 matrix (10000 x 10000 Float32, k = 16, 64 stacked restarts): restart-iterations/s and, with
NMFK_TILED_TIMING=1, the per-phase device time.  usage: fro_c3.py [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
X = synth.mixture(10000, 10000, 16, seed=2015, dtype=np.float32)
with nb.Context(0) as ctx:
    ctx.set_X(X)
    for rep in range(3):
        b = ctx.batch(16, 64)
        b.init_random(2015)
        ctx.solve([b], nb.default_params(maxiter=iters, variant=1))
        g = b.get(factors=False)
        print("solve_ms %.2f restart-iterations/s %.0f obj_norm[0] %.4f" % (ctx.last_solve_ms, g["iters"].sum() / ctx.last_solve_ms * 1e3, g["obj_norm"][0]), flush=True)
        b.close()
