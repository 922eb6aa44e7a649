import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo/nmfk.jl_b200/python")
import numpy as np, torch
import nmfk_b200 as nb
from nmfk_b200 import synth
n = m = 10000; k, R = 16, 64
X = synth.mixture(n, m, 16, seed=2015, dtype=np.float32)
Xpin = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
W0, H0 = synth.philox_inits(7, R, n, k, m, dtype=np.float32)
Wpin, Hpin = torch.from_numpy(W0).pin_memory(), torch.from_numpy(H0).pin_memory()
Wb, Hb = np.empty((k, n), np.float32), np.empty((m, k), np.float32)
with nb.Context(0) as ctx:
    params = nb.default_params(maxiter=2, engine=2)
    for i in range(2):
        ctx.set_X(Xpin.numpy().T)
        phi, rob, aic, tot = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        nb._lib.check(ctx._lib.nmfk_execute_run(ctx._h, k, R, C.c_void_p(Wpin.data_ptr()), C.c_void_p(Hpin.data_ptr()), 1, C.byref(params), Wb.ctypes.data_as(C.c_void_p), Hb.ctypes.data_as(C.c_void_p), C.byref(phi), C.byref(rob), C.byref(aic), C.byref(tot)), ctx._h)
