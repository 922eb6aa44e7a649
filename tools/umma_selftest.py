"""Runs the tcgen05 building-block self-test (nmfk_umma_selftest) one mode per process (a trapped
kernel poisons the CUDA context) and compares with NumPy.  usage: umma_selftest.py [mode]"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))


def one(mode):
    import numpy as np
    import nmfk_b200 as nb
    rng = np.random.default_rng(7)
    # tf32-exact inputs (multiples of 1/64 below 4) so that layout errors are not hidden by rounding
    U = (rng.integers(0, 256, (128, 16)) / 64.0).astype(np.float32)
    V = (rng.integers(0, 256, (64, 16)) / 64.0).astype(np.float32)
    Pss = np.zeros((128, 64), np.float32)
    Pts = np.zeros((128, 64), np.float32)
    Aa = np.zeros((128, 16), np.float32)
    Ab = np.zeros((128, 16), np.float32)
    err = C.c_int32(0)
    P = U.astype(np.float64) @ V.astype(np.float64).T
    ACC = (0.5 * P) @ V.astype(np.float64)
    with nb.Context(0) as ctx:
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        nb._lib.check(ctx._lib.nmfk_umma_selftest(ctx._h, p(U), p(V), mode, p(Pss), p(Pts), p(Aa), p(Ab), C.byref(err)), ctx._h)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    print(json.dumps(dict(mode=mode, err=err.value, P_ss=rel(Pss, P), P_ts=rel(Pts, P), ACC_a=rel(Aa, ACC), ACC_b=rel(Ab, ACC))))


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "--timing"):
    if len(sys.argv) > 1:
        one(int(sys.argv[1]))
    else:
        for mode in (1, 2, 4, 8, 24, 7):
            r = subprocess.run([sys.executable, __file__, str(mode)], capture_output=True, text=True, timeout=120)
            print(r.stdout.strip() or json.dumps(dict(mode=mode, failed=r.stderr.strip()[-300:])), flush=True)


def timing(reps=64):
    import numpy as np
    import nmfk_b200 as nb
    rng = np.random.default_rng(7)
    U = rng.random((128, 16)).astype(np.float32)
    U = (U.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)  # tf32-exact so that only the accumulation rounds
    V = rng.random((64, 16)).astype(np.float32)
    V = (V.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
    cyc = np.zeros(8, np.int64)
    acc = np.zeros((128, 16), np.float32)
    with nb.Context(0) as ctx:
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        nb._lib.check(ctx._lib.nmfk_umma_timing(ctx._h, p(U), p(V), reps, p(cyc), p(acc)), ctx._h)
    Q = np.tile(U, (1, 4)).astype(np.float64)  # 128 x 64
    exact = reps * (Q @ V.astype(np.float64))
    rel = (acc - exact) / exact
    names = ["1 MMA N=64 + commit + wait", "96 MMA N=64 (A tmem)", "96 MMA N=16 (A tmem)", "96 MMA N=32 (A tmem)",
             "16 x (ld x32 + wait)", "16 x (st x32 + wait)", "96 MMA N=64 (A smem)", "96 MMA N=16 alternating accumulators"]
    print(json.dumps(dict(cycles=dict(zip(names, cyc.tolist())), reps=reps, ksteps=reps * 8,
                          acc_rel_err_mean=float(rel.mean()), acc_rel_err_min=float(rel.min()), acc_rel_err_max=float(rel.max()))))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "--timing":
    timing(int(sys.argv[2]) if len(sys.argv) > 2 else 64)
