"""Variant FRO (method=:nmf, algorithm=:multdiv -> NMF.MultUpdate(obj=:mse); the update BASELINE.json's north star writes out) on
the GPU against the oracle's restatement of NMF.jl (oracle/nmfk_oracle.py::nmf_multupdate_mse):
the stacked-restart GEMM alone (tcgen05 3xTF32 / DMMA) against NumPy, then per-iteration factors, the stop rule, and execute
with method="nmf".  Tolerances: 1e-9 relative for Float64, 1e-4 for Float32 (north star)."""
import numpy as np
import pytest

import nmfk_b200 as nb
from nmfk_b200 import synth
from oracle import nmfk_oracle as o

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context()
    yield c
    c.close()


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 256), (256, 512, 1024), (1024, 1000, 2000), (48, 100, 36), (300, 260, 4100)])
def test_stacked_gemm_f32_3xtf32(ctx, M, N, K):
    """C = A B^T on tcgen05 with the 3-term TF32 split: FP32-level accuracy (plain TF32 would be ~1e-3), edge tiles in M, N, K."""
    rng = np.random.default_rng(M + N + K)
    A = rng.random((M, K), dtype=np.float32)
    B = rng.random((N, K), dtype=np.float32)
    C, ms = ctx.gemm_nt(A, B)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert relerr(C, ref) < 1e-5, relerr(C, ref)  # round-toward-zero chains of 96 instructions: ~ -5e-6
    # exactly representable inputs (TF32-exact, small K): the product is exact
    Ai = rng.integers(0, 64, (M, K)).astype(np.float32)
    Bi = rng.integers(0, 64, (N, K)).astype(np.float32)
    Ci, _ = ctx.gemm_nt(Ai[:, :32].copy(), Bi[:, :32].copy())
    assert np.array_equal(Ci.astype(np.float64), Ai[:, :32].astype(np.float64) @ Bi[:, :32].astype(np.float64).T)


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (256, 384, 1000), (100, 77, 333), (512, 2000, 1001)])
def test_stacked_gemm_f64_dmma(ctx, M, N, K):
    rng = np.random.default_rng(M * 3 + N + K)
    A = rng.random((M, K))
    B = rng.random((N, K))
    C, ms = ctx.gemm_nt(A, B)
    assert relerr(C, A @ B.T) < 1e-13


def _trace_fro(ctx, X, k, W0, H0, niter, dt):
    ctx.set_X(X)
    b = ctx.batch(k, W0.shape[0])
    b.set_init(W0, H0)
    outs = []
    for t in range(1, niter + 1):
        ctx.solve([b], nb.default_params(variant=1, maxiter=10 ** 6, iter_limit=t, normalize=0))
        g = b.get()
        outs.append((g["W"].copy(), g["H"].copy()))
    b.close()
    return outs


@pytest.mark.parametrize("n,m,k,R,niter,dt", [(512, 256, 8, 3, 12, np.float32), (1000, 200, 10, 2, 10, np.float64),
                                              (300, 64, 13, 2, 8, np.float64), (2048, 1024, 16, 5, 8, np.float32),
                                              (128, 64, 1, 2, 6, np.float32), (260, 132, 3, 2, 10, np.float32)])
def test_fro_per_iteration_vs_oracle(ctx, n, m, k, R, niter, dt):
    """Per-iteration W, H of the stacked FRO solve against the Float64 oracle from the same initial factors."""
    X = synth.mixture(n, m, 3, seed=17, dtype=dt)
    W0, H0 = synth.philox_inits(23, R, n, k, m, dtype=dt)
    outs = _trace_fro(ctx, X, k, W0, H0, niter, dt)
    tol = 1e-9 if dt == np.float64 else 1e-4
    delta = float(np.sqrt(np.finfo(dt).eps))
    X64 = X.astype(np.float64)
    for r in range(R):
        ref = []
        o.nmf_multupdate_mse(X64, k, Winit=W0[r].astype(np.float64), Hinit=H0[r].astype(np.float64), maxiter=niter, tol=0.0, delta=delta,
                             trace=lambda it, W, H: ref.append((W.copy(), H.copy())))
        for t in range(niter):
            assert relerr(outs[t][0][r], ref[t][0]) < tol, ("W", r, t, relerr(outs[t][0][r], ref[t][0]))
            assert relerr(outs[t][1][r], ref[t][1]) < tol, ("H", r, t, relerr(outs[t][1][r], ref[t][1]))


def test_fro_stop_rule_and_execute(ctx):
    """NMF.jl stop_condition (relative change of every column of W / row of H below tol) with a loose tol: same iteration count as
    the oracle; then execute(X, ks, nNMF; method=:nmf, algorithm=:multdiv): objective = normnan(X - W*H), rows of H sum to one,
    kopt and robustness against the oracle's execute_run fed with the same solver."""
    X = synth.mixture(400, 120, 3, seed=5)
    k, R = 3, 4
    W0, H0 = synth.philox_inits(9, R, 400, k, 120)
    ctx.set_X(X)
    b = ctx.batch(k, R)
    b.set_init(W0, H0)
    ctx.solve([b], nb.default_params(variant=1, maxiter=5000, tol=1e-4))
    g = b.get()
    b.close()
    for r in range(R):
        inf = {}
        W, H, obj = o.execute_singlerun_nmf(X.copy(), k, Winit=W0[r].copy(), Hinit=H0[r].copy(), maxiter=5000, tol=1e-4, info=inf)
        assert g["iters"][r] == inf["iters"] and g["stop_reason"][r] == 2, (r, g["iters"][r], inf)
        assert relerr(g["W"][r], W) < 1e-7 and relerr(g["H"][r], H) < 1e-7 and abs(g["obj_norm"][r] - obj) <= 1e-7 * obj + 1e-12
        assert np.allclose(g["H"][r].sum(axis=1), 1.0, atol=1e-12)
    # through the reference-shaped entry with the reference's keywords
    Wg, Hg, fg, rg, ag = nb.execute_run(X, k, R, inits=(W0, H0), method="nmf", algorithm="multdiv", maxiter=300, ctx=ctx)
    sols = [o.execute_singlerun_nmf(X.copy(), k, Winit=W0[r].copy(), Hinit=H0[r].copy(), maxiter=300) for r in range(R)]
    best = int(np.argmin([s[2] for s in sols]))
    assert abs(fg - sols[best][2]) <= 1e-7 * sols[best][2]
    labels, _ = o.clustersolutions([sols[i][1] for i in np.argsort([s[2] for s in sols], kind="stable")], False)
    _, _, csil, _, _ = o.finalize([s[0] for s in sols], [sols[i][1] for i in np.argsort([s[2] for s in sols], kind="stable")], labels, False)
    assert abs(rg - float(csil.min())) < 1e-7
    with pytest.raises(nb.NMFkError):
        nb.execute_run(X, k, R, method="nmf", algorithm="alspgrad", ctx=ctx)
