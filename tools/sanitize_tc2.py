"""compute-sanitizer case for the second-generation tcgen05 pass, the wide clustering walk and the sliced residual kernel:
ragged Float32 problem through execute_run with the tiled engine."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

X = synth.mixture(1156, 840, 4, seed=31, dtype=np.float32)
with nb.Context(0) as ctx:
    for k, R in ((12, 7), (5, 4), (16, 3)):
        W, H, phi, rob, aic = nb.execute_run(X, k, R, seed=3, ctx=ctx, maxiter=12, engine=2)
        print("k", k, "phi", phi, "rob", rob)
print("sanitize tc2 done")
