"""Short C3 run for ncu: 3 iterations of the 64 restarts through the tiled tcgen05 engine (6 tc_pass_kernel launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

n, m, k, R = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (10000, 10000, 16, 64)
dt = np.float64 if (len(sys.argv) > 5 and sys.argv[5] == "f64") else np.float32
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
X = synth.mixture(n, m, min(k, 16), seed=2015, dtype=dt)
with nb.Context(0) as ctx:
    ctx.set_X(X)
    b = ctx.batch(k, R)
    b.init_random(2015)
    ctx.solve([b], nb.default_params(maxiter=iters, engine=2))
    print("solve_ms", ctx.last_solve_ms, "iters", int(b.get(factors=False)["iters"].sum()))
    b.close()
