"""Stacked-restart GEMM of Variant FRO alone (nmfk_gemm_nt): correctness probe + TFLOP/s at the C3 / C4 stacked shapes.
usage: fro_gemm_bench.py [M N K dtype reps] ..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402

cases = [a.split(":") for a in sys.argv[1:]] or [["1024", "10000", "10000", "f32", "5"], ["8192", "2000", "20000", "f64", "2"]]
with nb.Context(0) as ctx:
    tf32 = max(ctx.measure_peak(13) for _ in range(2))
    dmma = max(ctx.measure_peak(1) for _ in range(2))
    for M, N, K, dt, reps in cases:
        M, N, K, reps = int(M), int(N), int(K), int(reps)
        dtype = np.float32 if dt == "f32" else np.float64
        rng = np.random.default_rng(1)
        A = rng.random((M, K), dtype=np.float32).astype(dtype)
        B = rng.random((N, K), dtype=np.float32).astype(dtype)
        C, ms = ctx.gemm_nt(A, B, reps=reps)
        sub = slice(0, min(M, 256))
        ref = A[sub].astype(np.float64) @ B.astype(np.float64).T
        err = float(np.max(np.abs(C[sub] - ref)) / np.max(np.abs(ref)))
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        peak = tf32 / 3 if dt == "f32" else dmma
        print(json.dumps(dict(M=M, N=N, K=K, dtype=dt, ms=ms, algorithmic_tflops=tf, peak=peak, frac=tf / peak, relerr_first_rows=err,
                              cluster=os.environ.get("NMFK_FRO_CLUSTER", "1"))), flush=True)
