"""Small end-to-end case for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

X = synth.mixture(70, 33, 3, seed=3)
Xn = X.copy()
Xn[5, 7] = np.nan
Xn[9, 1] = 0.0
with nb.Context(0) as ctx:
    for data in (X, Xn, X.astype(np.float32)):
        W, H, fit, rob, aic, kopt = nb.execute(data, range(1, 5), 5, seed=1, ctx=ctx, maxiter=60)
        print("kopt", kopt, "rob", rob)
print("sanitize case done")
