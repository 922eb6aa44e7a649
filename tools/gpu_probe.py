"""First-contact probe on the B200: roofline denominators + timing of the resident engine on C2.
Writes one JSON document to stdout (gpurun_out/probe.json)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

out = {}
try:
    out["nvidia_smi"] = subprocess.run(
        ["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
        capture_output=True, text=True).stdout.strip()
except Exception as e:  # pragma: no cover
    out["nvidia_smi"] = repr(e)

ctx = nb.Context(0)
peaks = {}
for name, which in (("fp64_dfma_tflops", 0), ("fp64_dmma_tflops", 1), ("fp32_ffma_tflops", 2), ("copy_gbs", 3),
                    ("lat_dfma_cyc", 4), ("lat_ffma_cyc", 5), ("lat_rcp64h_dfma_cyc", 6)):
    peaks[name] = [ctx.measure_peak(which) for _ in range(2)]
out["peaks"] = peaks

X = synth.mixture(1000, 200, 5, seed=2015)
ctx.set_X(X)
n, m = X.shape


def run(ks, R, **pk):
    bs = [ctx.batch(k, R) for k in ks]
    for b in bs:
        b.init_random(2015)
    t0 = time.perf_counter()
    ctx.solve(bs, nb.default_params(**pk))
    wall = time.perf_counter() - t0
    ms = ctx.last_solve_ms
    iters = {k: b.get(factors=False)["iters"] for k, b in zip(ks, bs)}
    for b in bs:
        b.close()
    tot = int(sum(int(v.sum()) for v in iters.values()))
    flops = sum(8.0 * n * m * k * float(v.sum()) for k, v in iters.items())
    return dict(ks=list(ks), R=R, ms=ms, wall_ms=wall * 1e3, total_iters=tot, restart_iters_per_s=tot / (ms * 1e-3),
                tflops=flops / (ms * 1e-3) / 1e12,
                iters_minmeanmax={k: [int(v.min()), float(v.mean()), int(v.max())] for k, v in iters.items()})


res = []
res.append(dict(tag="warmup k=2 R=8 100it", **run([2], 8, maxiter=100)))
for k in (2, 4, 6, 8, 10):
    res.append(dict(tag="fixed100 k=%d R=296" % k, **run([k], 296, maxiter=100)))
res.append(dict(tag="fixed100 k=10 R=100", **run([10], 100, maxiter=100)))
for k in (2, 6, 10):
    res.append(dict(tag="SCALAR fixed100 k=%d R=148" % k, **run([k], 148, maxiter=100, engine=3)))
    res.append(dict(tag="DMMA   fixed100 k=%d R=148" % k, **run([k], 148, maxiter=100, engine=1)))
res.append(dict(tag="fixed100 k=10 R=148", **run([10], 148, maxiter=100)))
res.append(dict(tag="fixed100 k=2:10 R=100", **run(range(2, 11), 100, maxiter=100)))
res.append(dict(tag="C2 full stop rule k=2:10 R=100", **run(range(2, 11), 100)))
res.append(dict(tag="C2 full stop rule k=2:10 R=100 (repeat)", **run(range(2, 11), 100)))
out["runs"] = res

# tiled engine on the same shape (engine=2) and on a mid-size shape it is meant for
res.append(dict(tag="TILED fixed100 k=10 R=100", **run([10], 100, maxiter=100, engine=2)))
res.append(dict(tag="TILED fixed100 k=4 R=100", **run([4], 100, maxiter=100, engine=2)))
out["runs"] = res
# end-to-end execute (solve + clustering + selection) through the public API
t0 = time.perf_counter()
det = {}
W, H, fit, rob, aic, kopt = nb.execute(X, range(2, 11), 100, seed=2015, ctx=ctx, details=det)
out["execute_c2"] = dict(wall_s=time.perf_counter() - t0, kopt=kopt, robustness=[float(x) for x in rob],
                         fit=[float(x) for x in fit], total_iters=det["total_iters"], solve_ms=det["solve_ms"])
# float32 variant of the same shape
X32 = X.astype(np.float32)
ctx.set_X(X32)
res32 = []
for k in (4, 10):
    res32.append(dict(tag="f32 fixed100 k=%d R=296" % k, **run([k], 296, maxiter=100)))
out["runs_f32"] = res32
# mid-size shapes for the tiled engine: f32 4000x4000 k=16 R=16 and f64 20000x1000 k=16 R=16
Xm = synth.mixture(4000, 4000, 16, seed=3, dtype=np.float32)
ctx.set_X(Xm)
n, m = Xm.shape
out["runs_tiled"] = [dict(tag="TILED f32 4000x4000 k=16 R=16 20it", **run([16], 16, maxiter=20))]
Xm = synth.mixture(20000, 1000, 8, seed=3)
ctx.set_X(Xm)
n, m = Xm.shape
out["runs_tiled"].append(dict(tag="TILED f64 20000x1000 k=16 R=16 20it", **run([16], 16, maxiter=20)))
out["runs_tiled"].append(dict(tag="TILED f64 20000x1000 k=32 R=32 20it", **run([32], 32, maxiter=20)))
ctx.close()
print(json.dumps(out, indent=1))
