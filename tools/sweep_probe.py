"""Does a multi-k solve pack as well as the per-k solves?  (fixed iteration counts, full waves)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import nmfk_b200 as nb
from nmfk_b200 import synth
X = synth.mixture(1000, 200, 5, seed=2015)
with nb.Context(0) as ctx:
    ctx.set_X(X)
    def run(ks, R, maxiter):
        bs = [ctx.batch(k, R) for k in ks]
        for b in bs: b.init_random(2015)
        ctx.solve(bs, nb.default_params(maxiter=maxiter))
        ms = ctx.last_solve_ms
        for b in bs: b.close()
        return ms
    run([2], 8, 50)
    for maxiter in (200, 400, 1000):
        solo = [run([k], 148, maxiter) for k in range(2, 11)]
        both = run(list(range(2, 11)), 148, maxiter)
        print(json.dumps(dict(maxiter=maxiter, sum_solo_ms=round(sum(solo), 2), sweep_ms=round(both, 2), solo=[round(v, 1) for v in solo])))
