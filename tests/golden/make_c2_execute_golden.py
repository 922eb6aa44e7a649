#!/usr/bin/env python3
"""Generates tests/golden/c2_execute_k4_6.npz: the CPU oracle's NMFk.execute(X, 4:6, 10; method=:simple) on the BASELINE C2
matrix (synthetic mixture 1000 x 200 Float64, k0 = 5, seed 2015) with the reference's full stop rule (maxiter = 10000), restart i
of every k started from Philox(key = 2015 + i).  ~25 s per restart on the build container, hence a committed fixture instead
of a live oracle run inside the GPU test (tests/test_gpu_configs.py::test_c2_execute_matches_oracle_golden).

    python tests/golden/make_c2_execute_golden.py [k ...]        # one process per k is fine; results are merged
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "nmfk.jl_b200", "python")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

from nmfk_b200 import synth  # noqa: E402
from oracle import nmfk_oracle as o  # noqa: E402

KS, R, SEED = (4, 5, 6), 10, 2015
HERE = os.path.dirname(os.path.abspath(__file__))


def one_k(k):
    X = synth.mixture(1000, 200, 5, seed=2015)
    det = {}
    W, H, fit, rob, aic = o.execute_k(X.copy(), k, R, seed=SEED, details=det)
    stop = {"maxiter": 1, "tol": 2, "reattempts": 3, "consistency": 4}
    out = {"k%d_iters" % k: det["iters"], "k%d_objvalue" % k: det["objvalue"], "k%d_idxsort" % k: det["idxsort"],
           "k%d_labels" % k: det["labels"], "k%d_clustersil" % k: det["clustersil"][:, 0], "k%d_fit" % k: fit,
           "k%d_rob" % k: rob, "k%d_aic" % k: aic, "k%d_W" % k: W, "k%d_H" % k: H,
           "k%d_stop" % k: np.asarray([stop[s] for s in det["stop_reasons"]])}
    np.savez(os.path.join(HERE, "_c2_part_k%d.npz" % k), **out)


if __name__ == "__main__":
    ks = [int(a) for a in sys.argv[1:]] or list(KS)
    for k in ks:
        one_k(k)
    parts = [os.path.join(HERE, "_c2_part_k%d.npz" % k) for k in KS]
    if all(os.path.exists(p) for p in parts):
        merged = {}
        for p in parts:
            merged.update(dict(np.load(p)))
        rob = [float(merged["k%d_rob" % k]) for k in KS]
        merged["kopt"] = np.asarray(-1 if o.getk(list(KS), rob) is None else o.getk(list(KS), rob))
        np.savez_compressed(os.path.join(HERE, "c2_execute_k4_6.npz"), **merged)
        for p in parts:
            os.remove(p)
        print("wrote c2_execute_k4_6.npz kopt=%s rob=%s" % (merged["kopt"], rob))
