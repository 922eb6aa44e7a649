"""Clustering phase (K10-K12: clustersolutions + pairwise cosine + silhouettes) alone at the C4 k = 32 size (R = 256 solutions of
k = 32 rows of length m = 2000 -> R*k = 8192 points, an 8192 x 8192 Float64 distance matrix = 537 MB) and at the size the 8-GPU
sweep of C2 clusters (R = 800, k = 10, m = 200).  usage: cluster_bench.py   (wrap in ncu --metrics gpu__time_duration.sum for the
per-kernel split)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402

rng = np.random.default_rng(0)
for R, k, m in ((256, 32, 2000), (800, 10, 200)):
    base = rng.random((k, m)) ** 3
    H = np.abs(np.stack([base[rng.permutation(k)] * (1 + 0.05 * rng.standard_normal((k, m))) for _ in range(R)]))
    obj = rng.random(R)
    with nb.Context(0) as ctx:
        ctx.set_X(np.ones((8, m)))
        b = ctx.import_solutions(H, obj)
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            cl = b.cluster()
            ts.append(time.perf_counter() - t0)
        b.close()
    N = R * k
    print(json.dumps(dict(R=R, k=k, m=m, points=N, wall_ms=[round(t * 1e3, 2) for t in ts], robustness=cl["robustness"],
                          dist_matrix_MB=N * N * 8 / 1e6, gram_gflop=2.0 * N * N * m / 1e9)), flush=True)
