"""Fixed-iteration timing of the tiled engine on the large configurations (C3 / C4 / C5 shapes, or
scaled-down versions of them) with the roofline fraction of SURVEY.md §8d (8nmk flops per
restart-iteration against the measured FP32 FFMA / FP64 DMMA / TF32 mma.sync peaks).

usage: tiled_bench.py [case ...]   case = name:n:m:k:R:dtype:iters   -> JSON lines on stdout
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

DEFAULT = [
    "C3:10000:10000:16:64:f32:10",
    "C4k32:100000:2000:32:32:f64:10",
    "C4k8:100000:2000:8:32:f64:10",
    "C5q:500000:1000:24:32:f32:10",
]


def main():
    cases = sys.argv[1:] or DEFAULT
    peaks = None
    for case in cases:
        name, n, m, k, R, dt, iters = case.split(":")
        n, m, k, R, iters = int(n), int(m), int(k), int(R), int(iters)
        dtype = np.float32 if dt == "f32" else np.float64
        X = synth.mixture(n, m, min(k, 16), seed=2015, dtype=dtype)
        with nb.Context(0) as ctx:
            if peaks is None:
                peaks = dict(ffma=max(ctx.measure_peak(2) for _ in range(2)), dmma=max(ctx.measure_peak(1) for _ in range(2)),
                             tf32_mma_sync=max(ctx.measure_peak(11) for _ in range(2)),
                             bf16_mma_sync=max(ctx.measure_peak(12) for _ in range(2)))
                print(json.dumps(dict(peaks_tflops=peaks)), flush=True)
            ctx.set_X(X)
            ms = None
            for engine_iters in (10, iters, iters):  # warm-up (with one check: loads every kernel), then best of two
                b = ctx.batch(k, R)
                b.init_random(2015)
                ctx.solve([b], nb.default_params(maxiter=engine_iters, engine=int(os.environ.get("NMFK_BENCH_ENGINE", "2"))))
                if engine_iters == iters and (ms is None or ctx.last_solve_ms < ms):
                    ms = ctx.last_solve_ms
                out = b.get(factors=False)
                b.close()
            tot = int(out["iters"].sum())
            flops = 8.0 * n * m * k * tot
            peak = peaks["ffma"] if dt == "f32" else peaks["dmma"]
            print(json.dumps(dict(case=name, n=n, m=m, k=k, R=R, dtype=dt, iters=tot, ms=round(ms, 3),
                                  rit_per_s=round(tot / ms * 1e3, 2), tflops=round(flops / ms / 1e9, 3),
                                  frac_of_fma_peak=round(flops / ms / 1e9 / peak, 4),
                                  obj_norm_first=float(out["obj_norm"][0]))), flush=True)


if __name__ == "__main__":
    main()
