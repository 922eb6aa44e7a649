"""Randomised shapes through the second-generation tcgen05 pass (k <= 16) against the scalar-FMA pass of the same engine: ragged
own / reduction sizes, 1..9 restarts (odd groups, shadow units), sliced reductions, frozen restarts.  usage: tc2_stress.py [cases]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(123)
worst = 0.0
with nb.Context(0) as ctx:
    for c in range(ncases):
        n = int(rng.choice([4, 8, 64, 128, 132, 260, 1000, 3108, 20000])) if c % 3 else 4 * int(rng.integers(1, 900))
        m = 4 * int(rng.integers(1, 700)) if c % 2 else int(rng.choice([4, 64, 68, 200, 332, 5000]))
        k = int(rng.integers(1, 17))
        R = int(rng.integers(1, 10))
        iters = int(rng.integers(1, 5))
        X = synth.mixture(n, m, min(k, 4), seed=c, dtype=np.float32)
        W0, H0 = synth.philox_inits(100 + c, R, n, k, m, dtype=np.float32)
        ctx.set_X(X)
        out = []
        for eng in (2, 4):
            b = ctx.batch(k, R)
            b.set_init(W0, H0)
            ctx.solve([b], nb.default_params(maxiter=iters, engine=eng))
            g = b.get()
            out.append((g["W"].astype(np.float64), g["H"].astype(np.float64)))
            b.close()
        e = max(np.max(np.abs(out[0][i] - out[1][i])) / max(np.max(np.abs(out[1][i])), 1e-300) for i in range(2))
        worst = max(worst, e)
        print("case %2d: n=%6d m=%5d k=%2d R=%d iters=%d  relerr %.2e" % (c, n, m, k, R, iters, e), flush=True)
        assert e < 1e-4, "mismatch"
print("tc2 stress OK, worst", worst)
