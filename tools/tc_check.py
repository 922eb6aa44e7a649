"""tcgen05 Float32 tiled engine (engine=2) against the scalar-FMA tiled engine (engine=4) and the Float64
oracle after a fixed number of iterations, several restarts per batch.  One case per process (a trapped
kernel poisons the CUDA context).  usage: tc_check.py [n:m:k:R:iters ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
sys.path.insert(0, ROOT)

DEFAULT = ["1024:256:16:4:5", "1000:200:10:5:12", "3108:332:7:9:12", "640:1204:24:7:6", "512:384:32:3:6", "256:128:3:1:20",
           "20000:1000:16:8:4"]


def one(case):
    import numpy as np
    import nmfk_b200 as nb
    from nmfk_b200 import synth
    n, m, k, R, iters = (int(v) for v in case.split(":"))
    X = synth.mixture(n, m, min(k, 5), seed=17, dtype=np.float32)
    W0, H0 = synth.philox_inits(23, R, n, k, m, dtype=np.float32)
    res = {}
    with nb.Context(0) as ctx:
        ctx.set_X(X)
        for eng in (4, 2):
            b = ctx.batch(k, R)
            b.set_init(W0, H0)
            ctx.solve([b], nb.default_params(engine=eng, maxiter=iters, normalize=0))
            res[eng] = b.get()
            res[eng]["ms"] = ctx.last_solve_ms
            b.close()
    rel = lambda a, b: float(np.max(np.abs(a.astype(np.float64) - b)) / np.max(np.abs(b)))
    out = dict(case=case, W_tc_vs_scalar=rel(res[2]["W"], res[4]["W"].astype(np.float64)),
               H_tc_vs_scalar=rel(res[2]["H"], res[4]["H"].astype(np.float64)), ms_scalar=round(res[4]["ms"], 3),
               ms_tc=round(res[2]["ms"], 3))
    if n * m <= 400000:
        from oracle import nmfk_oracle as o
        Wo, Ho = [], []
        for r in range(min(R, 2)):
            Xc = np.array(X, dtype=np.float64, order="F", copy=True)
            W, H, _ = o.nmf_multiplicative(Xc, k, Winit=W0[r].astype(np.float64), Hinit=H0[r].astype(np.float64), maxiter=iters)
            out["W_tc_vs_oracle_r%d" % r] = rel(res[2]["W"][r], W)
            out["H_tc_vs_oracle_r%d" % r] = rel(res[2]["H"][r], H)
            out["W_scalar_vs_oracle_r%d" % r] = rel(res[4]["W"][r], W)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        for case in (sys.argv[1:] or DEFAULT):
            try:
                r = subprocess.run([sys.executable, __file__, "--one", case], capture_output=True, text=True, timeout=180)
                print(r.stdout.strip() or json.dumps(dict(case=case, failed=r.stderr.strip()[-400:])), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps(dict(case=case, failed="timeout")), flush=True)
