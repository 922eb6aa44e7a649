"""Per-k timing of the resident engine on the C2 workload (which k dominates the sweep, and how much
of a step is tail).  usage: perk.py [R]   -> JSON lines on stdout"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100
X = synth.mixture(1000, 200, 5, seed=2015)
n, m = X.shape
with nb.Context(0) as ctx:
    ctx.set_X(X)
    peak = max(ctx.measure_peak(1) for _ in range(2))

    def run(ks, Rr, **pk):
        bs = [ctx.batch(k, Rr) for k in ks]
        for b in bs:
            b.init_random(2015)
        ctx.solve(bs, nb.default_params(**pk))
        ms = ctx.last_solve_ms
        its = {k: b.get(factors=False)["iters"] for k, b in zip(ks, bs)}
        for b in bs:
            b.close()
        flops = sum(8.0 * n * m * k * float(v.sum()) for k, v in its.items())
        tot = sum(int(v.sum()) for v in its.values())
        return dict(ks=list(ks), R=Rr, ms=round(ms, 3), iters=tot, max_iters=max(int(v.max()) for v in its.values()),
                    rit_per_s=round(tot / ms * 1e3), tflops=round(flops / ms / 1e9, 3), frac=round(flops / ms / 1e9 / peak, 4))

    run([2], 8, maxiter=50)
    print(json.dumps(dict(peak_dmma_tflops=peak)))
    # steady state: 148 restarts x 200 iterations, no stop rule interference (full waves)
    for k in range(2, 11):
        print(json.dumps(dict(tag="fixed200 R=148", **run([k], 148, maxiter=200))))
    for k in (4, 10):
        print(json.dumps(dict(tag="fixed200 R=148 SCALAR", **run([k], 148, maxiter=200, engine=3))))
    # the real stop rule, one k at a time, then the whole sweep
    for k in range(2, 11):
        print(json.dumps(dict(tag="stoprule", **run([k], R))))
    print(json.dumps(dict(tag="sweep", **run(list(range(2, 11)), R))))
