"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded initialisations.  Tolerances (BASELINE.json north_star): per-iteration W, H and fit
within 1e-9 relative for Float64 and 1e-4 for Float32 over a fixed iteration count;
cluster assignments, robustness ranking and kopt identical."""
import numpy as np
import pytest

import nmfk_b200 as nb
from nmfk_b200 import synth
from oracle import nmfk_oracle as o

pytestmark = pytest.mark.gpu

RTOL64 = 1e-9
RTOL32 = 1e-4


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def oracle_trace(X, k, W0, H0, niter, **kw):
    Ws, Hs, objs = [], [], []

    def cb(it, W, H, obj):
        Ws.append(W.copy())
        Hs.append(H.copy())

    Xc = np.array(X, dtype=np.float64, order="F", copy=True)
    o.nmf_multiplicative(Xc, k, Winit=W0.astype(np.float64), Hinit=H0.astype(np.float64), maxiter=niter, trace=cb, **kw)
    # the :74 objective of every iteration's W,H against the lambda-substituted X
    Xl = np.array(X, dtype=np.float64, copy=True)
    inan = np.isnan(Xl)
    Xl[Xl <= 0] = 1e-32
    for W, H in zip(Ws, Hs):
        E = (Xl - W @ H)[~inan]
        objs.append(float(np.sum(E ** 2)))
    return np.stack(Ws), np.stack(Hs), np.asarray(objs)


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context()
    yield c
    c.close()


@pytest.mark.parametrize("n,m,k,niter", [(15, 5, 2, 60), (15, 5, 3, 60), (15, 5, 5, 45), (50, 20, 4, 40), (37, 300, 7, 25),
                                         (1000, 200, 10, 25), (300, 64, 13, 22), (100, 40, 32, 22), (64, 9, 1, 30)])
def test_trace_parity_f64(ctx, n, m, k, niter):
    X = synth.readme_bss() if (n, m) == (15, 5) else synth.mixture(n, m, 3, seed=7)
    W0, H0 = synth.philox_inits(11, 1, n, k, m)
    Wt, Ht, ob = nb.trace(X, k, W0[0], H0[0], niter, ctx=ctx)
    Wr, Hr, obr = oracle_trace(X, k, W0[0], H0[0], niter)
    assert len(Wr) == niter, "oracle stopped early; pick another case"
    for t in range(niter):
        assert relerr(Wt[t], Wr[t]) < RTOL64, ("W", t)
        assert relerr(Ht[t], Hr[t]) < RTOL64, ("H", t)
    assert np.allclose(ob, obr, rtol=1e-7, atol=1e-18 + 1e-9 * obr.max())


@pytest.mark.parametrize("n,m,k,niter", [(15, 5, 3, 40), (200, 50, 6, 30), (1000, 200, 10, 20), (128, 96, 16, 20)])
def test_trace_parity_f32(ctx, n, m, k, niter):
    X = synth.mixture(n, m, 4, seed=5, dtype=np.float32)
    W0, H0 = synth.philox_inits(3, 1, n, k, m, dtype=np.float32)
    Wt, Ht, ob = nb.trace(X, k, W0[0], H0[0], niter, ctx=ctx)
    Wr, Hr, obr = oracle_trace(X, k, W0[0], H0[0], niter)  # the reference computes in Float64
    for t in range(niter):
        assert relerr(Wt[t], Wr[t]) < RTOL32, ("W", t)
        assert relerr(Ht[t], Hr[t]) < RTOL32, ("H", t)


@pytest.mark.parametrize("n,m,k,niter", [(15, 5, 3, 40), (1000, 200, 10, 22), (203, 77, 6, 22), (300, 64, 13, 22),
                                         (90, 40, 32, 15), (64, 9, 1, 20)])
def test_trace_parity_f64_scalar_formulation(ctx, n, m, k, niter):
    """engine=3: the scalar-FMA formulation of the resident engine (the default for Float64 is DMMA)."""
    X = synth.mixture(n, m, 3, seed=7)
    W0, H0 = synth.philox_inits(11, 1, n, k, m)
    Wt, Ht, ob = nb.trace(X, k, W0[0], H0[0], niter, ctx=ctx, engine=3)
    Wr, Hr, obr = oracle_trace(X, k, W0[0], H0[0], niter)
    for t in range(niter):
        assert relerr(Wt[t], Wr[t]) < RTOL64 and relerr(Ht[t], Hr[t]) < RTOL64, t


def test_full_stop_rule_matches_oracle(ctx):
    """Reference stop rule (tolOF/baditers/reattempts state machine): same iteration counts,
    stop reasons, objective and normalised factors for every restart."""
    X = synth.readme_bss()
    for k in (2, 3, 4):
        R = 6
        W0, H0 = synth.philox_inits(100, R, 15, k, 5)
        ctx.set_X(X)
        b = ctx.batch(k, R)
        b.set_init(W0, H0)
        ctx.solve([b])
        g = b.get()
        for r in range(R):
            inf = {}
            Xc = X.copy()
            W, H, of = o.execute_singlerun_compute(Xc, k, Winit=W0[r].copy(), Hinit=H0[r].copy(), info=inf)
            assert np.array_equal(Xc, X)
            assert g["iters"][r] == inf["iters"], (k, r)
            assert g["stop_reason"][r] == {"maxiter": 1, "tol": 2, "reattempts": 3, "consistency": 4}[inf["stop_reason"]]
            assert g["obj_norm"][r] == pytest.approx(of, rel=1e-6, abs=1e-12)
            assert relerr(g["W"][r], W) < 1e-7 and relerr(g["H"][r], H) < 1e-7
            assert np.allclose(g["H"][r].sum(axis=1), 1.0, atol=1e-12)  # test_execute_smoke.jl:18-19
        b.close()


def test_execute_matches_oracle_decisions(ctx):
    """execute(X, 2:5, 10): kopt, cluster labels, robustness, fit, aic vs the oracle from the same
    Philox streams (device-generated here, NumPy-generated there)."""
    X = synth.readme_bss()
    det = {}
    W, H, fit, rob, aic, kopt = nb.execute(X, range(2, 6), 10, seed=100, ctx=ctx, details=det)
    Wo, Ho, fito, robo, aico, kopto = o.execute(X.copy(), range(2, 6), 10, seed=100)
    assert kopt == kopto == 3
    assert fit[0] == np.inf and rob[0] == -1
    for k in range(2, 6):
        assert rob[k - 1] == pytest.approx(robo[k - 1], abs=1e-6), k
        assert fit[k - 1] == pytest.approx(fito[k - 1], rel=1e-5, abs=1e-10), k
        assert aic[k - 1] == pytest.approx(aico[k - 1], rel=1e-5), k
        assert relerr(W[k], Wo[k]) < 1e-6 and relerr(H[k], Ho[k]) < 1e-6
    assert list(np.argsort(-rob[1:5], kind="stable")) == list(np.argsort(-robo[1:5], kind="stable"))
    assert det["total_iters"] > 0


def test_cluster_labels_bit_identical(ctx):
    """clustersolutions + silhouettes on the solver's H stacks: labels identical to the oracle,
    silhouettes to 1e-9."""
    X = synth.mixture(60, 24, 3, seed=9)
    for k in (2, 3, 5):
        R = 12
        ctx.set_X(X)
        b = ctx.batch(k, R)
        b.init_random(500)
        ctx.solve([b], nb.default_params(maxiter=200))
        g = b.get()
        cl = b.cluster()
        order = np.argsort(g["obj_norm"], kind="stable")
        assert list(cl["order"]) == list(order)
        Hs = [np.array(g["H"][i], dtype=np.float64) for i in order]
        Ws = [np.array(g["W"][i], dtype=np.float64) for i in order]
        labels, cent = o.clustersolutions(Hs, False)
        assert np.array_equal(cl["labels"], labels)
        _, _, csil, _, _ = o.finalize(Ws, Hs, labels, False)
        assert np.allclose(cl["clustersil"], csil[:, 0], atol=1e-9)
        assert cl["robustness"] == pytest.approx(float(csil.min()), abs=1e-9)
        assert np.allclose(cl["centroids"], cent, rtol=1e-10)
        b.close()


def test_cluster_unit_case_from_reference(ctx):
    """test/test_cluster_unit.jl:36-54 through the device path (clusterWmatrix=true)."""
    f1 = np.array([[1.0, 0], [0, 1], [1, 0], [0, 1]])
    f2 = np.array([[0.0, 1], [1, 0], [0, 1], [1, 0]])
    ctx.set_X(np.ones((4, 3)))
    b = ctx.batch(2, 2)
    H = np.ones((2, 2, 3))
    b.set_init(np.stack([f1, f2]), H)
    ctx.solve([b], nb.default_params(maxiter=0, normalize=0))  # no iterations: just evaluate + keep factors
    g = b.get()
    assert np.array_equal(g["W"][0], f1) and np.array_equal(g["W"][1], f2)
    cl = b.cluster(clusterWmatrix=True)
    assert cl["labels"].shape == (2, 2)
    assert list(cl["labels"][:, 0]) == [1, 2] and list(cl["labels"][:, 1]) == [2, 1]
    assert cl["centroids"].shape == (2, 4)
    b.close()


def test_device_philox_init_equals_numpy(ctx):
    X = synth.mixture(33, 17, 2, seed=1)
    ctx.set_X(X)
    b = ctx.batch(3, 4)
    b.init_random(41)
    ctx.solve([b], nb.default_params(maxiter=0, normalize=0))
    g = b.get()
    W0, H0 = synth.philox_inits(41, 4, 33, 3, 17)
    assert np.array_equal(g["W"], W0) and np.array_equal(g["H"], H0)
    b.close()


def test_zeros_nan_and_errors(ctx):
    rng = np.random.default_rng(3)
    X = synth.mixture(40, 12, 3, seed=2)
    X[rng.random(X.shape) < 0.1] = 0.0
    xi = ctx.set_X(X)
    assert xi.nzero == int((X <= 0).sum()) and xi.nnan == 0
    W0, H0 = synth.philox_inits(5, 1, 40, 3, 12)
    Wt, Ht, ob = nb.trace(X, 3, W0[0], H0[0], 30, ctx=ctx)
    Wr, Hr, obr = oracle_trace(X, 3, W0[0], H0[0], 30)
    assert relerr(Wt[-1], Wr[-1]) < RTOL64 and relerr(Ht[-1], Hr[-1]) < RTOL64
    # NaN entries: EM-style imputation X[inan] = (W*H)[inan] (NMFkMultiplicative.jl:72; runtests.jl:270-273)
    Xn = synth.mixture(40, 12, 3, seed=2)
    Xn[rng.random(Xn.shape) < 0.2] = np.nan
    xi = ctx.set_X(Xn)
    assert xi.nnan == int(np.isnan(Xn).sum())
    Wt, Ht, ob = nb.trace(Xn, 3, W0[0], H0[0], 35, ctx=ctx)
    Wr, Hr, obr = oracle_trace(Xn, 3, W0[0], H0[0], 35)
    for t in (0, 1, 9, 10, 34):
        assert relerr(Wt[t], Wr[t]) < RTOL64 and relerr(Ht[t], Hr[t]) < RTOL64, t
    # negative entries throw (NMFkMultiplicative.jl:4-7) ...
    Xneg = X.copy()
    Xneg[3, 4] = -1.0
    with pytest.raises(nb.NegativeEntriesError):
        ctx.set_X(Xneg)
    # ... unless a NaN hides them from minimum()
    Xneg[0, 0] = np.nan
    ctx.set_X(Xneg)
    # NaN initial values are an error (:42-44)
    ctx.set_X(X)
    b = ctx.batch(3, 1)
    Wbad = W0.copy()
    Wbad[0, 2, 1] = np.nan
    with pytest.raises(nb.NMFkError) as ei:
        b.set_init(Wbad, H0)
    assert ei.value.status == -3
    b.close()


def test_fixed_factors(ctx):
    """Wfixed / Hfixed (NMFkMultiplicative.jl:66,69): the fixed factor never changes."""
    X = synth.mixture(30, 10, 2, seed=4)
    W0, H0 = synth.philox_inits(8, 1, 30, 2, 10)
    for kw in ({"Wfixed": True}, {"Hfixed": True}):
        Wt, Ht, ob = nb.trace(X, 2, W0[0], H0[0], 25, ctx=ctx, **kw)
        Wr, Hr, obr = oracle_trace(X, 2, W0[0], H0[0], 25, **kw)
        assert relerr(Wt[-1], Wr[-1]) < RTOL64 and relerr(Ht[-1], Hr[-1]) < RTOL64


def test_c2_shape_sweep_properties(ctx):
    """BASELINE config C2 shape at reduced restart count: all k concurrently, reference stop rule.
    Size-independent properties: rows of H sum to one, W,H >= 0, objective == ||X - WH||,
    rank-5 data => fit collapses at k >= 5, robustness high at the true rank."""
    X = synth.mixture(1000, 200, 5, seed=2015)
    ctx.set_X(X)
    ks = list(range(2, 11))
    bs = [ctx.batch(k, 8) for k in ks]
    for b in bs:
        b.init_random(2015)
    ctx.solve(bs)
    for k, b in zip(ks, bs):
        g = b.get()
        assert (g["W"] >= 0).all() and (g["H"] >= 0).all()
        assert np.allclose(g["H"].sum(axis=2), 1.0, atol=1e-10)
        r = int(np.argmin(g["obj_norm"]))
        assert np.linalg.norm(X - g["W"][r] @ g["H"][r]) == pytest.approx(g["obj_norm"][r], rel=1e-8)
        assert (g["iters"] % 10 == 0).all() and (g["iters"] <= 10000).all()
        assert set(g["stop_reason"]) <= {1, 2, 3, 4}
        if k >= 5:
            assert g["obj_norm"].min() < 1.0
        b.close()


# ---------------------------------------------------------------------------------------------
# tiled engine (factors streamed; BASELINE configs C3-C5), forced on small shapes via engine=2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m,k,niter,dt", [(1000, 200, 10, 22, np.float64), (300, 64, 13, 22, np.float64),
                                            (5000, 40, 4, 21, np.float64), (40, 3000, 3, 21, np.float64),
                                            (700, 300, 32, 12, np.float64), (777, 333, 7, 21, np.float32),
                                            (15, 5, 3, 40, np.float64)])
def test_tiled_trace_parity(ctx, n, m, k, niter, dt):
    X = synth.mixture(n, m, 3, seed=17, dtype=dt)
    W0, H0 = synth.philox_inits(23, 1, n, k, m, dtype=dt)
    Wt, Ht, ob = nb.trace(X, k, W0[0], H0[0], niter, ctx=ctx, engine=2)
    Wr, Hr, obr = oracle_trace(X, k, W0[0], H0[0], niter)
    tol = RTOL64 if dt == np.float64 else RTOL32
    for t in range(niter):
        assert relerr(Wt[t], Wr[t]) < tol, ("W", t)
        assert relerr(Ht[t], Hr[t]) < tol, ("H", t)
    if dt == np.float64:
        assert np.allclose(ob, obr, rtol=1e-7, atol=1e-18 + 1e-9 * obr.max())


def test_tiled_nan_and_fixed(ctx):
    rng = np.random.default_rng(5)
    Xn = synth.mixture(300, 70, 3, seed=2)
    Xn[rng.random(Xn.shape) < 0.15] = np.nan
    Xn[rng.random(Xn.shape) < 0.05] = 0.0
    W0, H0 = synth.philox_inits(5, 1, 300, 3, 70)
    Wt, Ht, ob = nb.trace(Xn, 3, W0[0], H0[0], 25, ctx=ctx, engine=2)
    Wr, Hr, obr = oracle_trace(Xn, 3, W0[0], H0[0], 25)
    for t in (0, 1, 9, 10, 24):
        assert relerr(Wt[t], Wr[t]) < RTOL64 and relerr(Ht[t], Hr[t]) < RTOL64, t
    X = synth.mixture(300, 70, 3, seed=2)
    for kw in ({"Wfixed": True}, {"Hfixed": True}):
        Wt, Ht, ob = nb.trace(X, 3, W0[0], H0[0], 15, ctx=ctx, engine=2, **kw)
        Wr, Hr, obr = oracle_trace(X, 3, W0[0], H0[0], 15, **kw)
        assert relerr(Wt[-1], Wr[-1]) < RTOL64 and relerr(Ht[-1], Hr[-1]) < RTOL64


def test_tiled_full_stop_rule_matches_resident_and_oracle(ctx):
    """Same restarts through both engines with the reference stop rule: identical iteration
    counts and stop reasons, factors equal to rounding; and equal to the oracle."""
    X = synth.readme_bss()
    k, R = 3, 6
    W0, H0 = synth.philox_inits(100, R, 15, k, 5)
    ctx.set_X(X)
    res = {}
    for eng in (1, 2, 3):
        b = ctx.batch(k, R)
        b.set_init(W0, H0)
        ctx.solve([b], nb.default_params(engine=eng))
        res[eng] = b.get()
        b.close()
    assert np.array_equal(res[1]["iters"], res[2]["iters"]) and np.array_equal(res[1]["iters"], res[3]["iters"])
    assert np.array_equal(res[1]["stop_reason"], res[2]["stop_reason"])
    assert relerr(res[3]["W"], res[1]["W"]) < 1e-8 and relerr(res[3]["H"], res[1]["H"]) < 1e-8
    assert relerr(res[2]["W"], res[1]["W"]) < 1e-8 and relerr(res[2]["H"], res[1]["H"]) < 1e-8
    assert np.allclose(res[2]["obj_norm"], res[1]["obj_norm"], rtol=1e-6, atol=1e-12)
    for r in range(R):
        inf = {}
        W, H, of = o.execute_singlerun_compute(X.copy(), k, Winit=W0[r].copy(), Hinit=H0[r].copy(), info=inf)
        assert res[2]["iters"][r] == inf["iters"]
        assert relerr(res[2]["W"][r], W) < 1e-7 and relerr(res[2]["H"][r], H) < 1e-7


# ---------------------------------------------------------------------------------------------
# row-sharded X code path (nmfk_ctx_comm_init): on one GPU with nranks == 1 the tiled engine runs
# the same kernels as the multi-GPU run (numerators to a buffer, "all-reduce", apply) without NCCL;
# tests/multigpu/rowshard_ranks.py is the 2-GPU run of the same thing under torchrun.
# ---------------------------------------------------------------------------------------------
def test_rowsharded_path_single_rank_trace_parity():
    n, m, k, niter = 300, 64, 5, 25
    X = synth.mixture(n, m, 3, seed=7)
    W0, H0 = synth.philox_inits(11, 1, n, k, m)
    with nb.Context() as c:
        c.comm_init(1, 0, None, 0, n)
        Wt, Ht, ob = nb.trace(X, k, W0[0], H0[0], niter, ctx=c, engine=2)
    Wr, Hr, obr = oracle_trace(X, k, W0[0], H0[0], niter)
    for t in range(niter):
        assert relerr(Wt[t], Wr[t]) < RTOL64, ("W", t)
        assert relerr(Ht[t], Hr[t]) < RTOL64, ("H", t)
    assert relerr(ob, obr) < 1e-8


def test_rowsharded_init_random_keeps_own_rows_of_the_global_stream():
    n_global, m, k, R = 400, 30, 4, 3
    r0, r1 = 100, 300
    X = synth.mixture(n_global, m, 3, seed=3)
    W0, H0 = synth.philox_inits(21, R, n_global, k, m)
    with nb.Context() as c:
        c.comm_init(1, 0, None, r0, n_global)
        c.set_X(X[r0:r1])
        b = c.batch(k, R)
        b.init_random(21)
        c.solve([b], nb.default_params(engine=2, maxiter=0, normalize=0))
        out = b.get()
        b.close()
    assert np.array_equal(out["W"], W0[:, r0:r1, :])
    assert np.array_equal(out["H"], H0)


def test_rowsharded_two_gpus_torchrun():
    """Real exchange over NCCL: needs two GPUs (skipped on a one-GPU box; run with gpurun --gpus 2)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(root, "tests", "multigpu", "rowshard_ranks.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ROWSHARD OK" in r.stdout


def test_restart_sharded_sweep_two_gpus_torchrun():
    """nmfk_sweep over NCCL on two GPUs against the oracle's execute with nNMF = 2 * R_local (skipped on a one-GPU box; the
    log of the builder's 2-GPU run is committed under profiles/)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613",
                        os.path.join(root, "tests", "multigpu", "sweep_ranks.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SWEEP OK" in r.stdout


# ---------------------------------------------------------------------------------------------
# Float32 tiled engine on the 5th-generation tensor cores (kl_tiled_tc.cu): tcgen05.mma kind::tf32 with the
# 3-term split, tensor-memory operands, per-unit accumulators drained in FP32.  engine=2 takes this path for
# Float32 data without NaN whose row and column counts are multiples of 4; engine=4 forces the scalar-FMA pass.
# ---------------------------------------------------------------------------------------------
def test_umma_building_blocks_selftest(ctx):
    """tcgen05.mma with A/B in shared memory, A in tensor memory, and the 2-term split of the A operand."""
    import ctypes as C
    rng = np.random.default_rng(7)
    U = (rng.integers(0, 256, (128, 16)) / 64.0).astype(np.float32)  # tf32-exact inputs
    V = (rng.integers(0, 256, (64, 16)) / 64.0).astype(np.float32)
    Pss, Pts = np.zeros((128, 64), np.float32), np.zeros((128, 64), np.float32)
    Aa, Ab = np.zeros((128, 16), np.float32), np.zeros((128, 16), np.float32)
    err = C.c_int32(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    nb._lib.check(ctx._lib.nmfk_umma_selftest(ctx._h, p(U), p(V), 7, p(Pss), p(Pts), p(Aa), p(Ab), C.byref(err)), ctx._h)
    P = U.astype(np.float64) @ V.astype(np.float64).T
    assert err.value == 0
    assert np.array_equal(Pss.astype(np.float64), P) and np.array_equal(Pts.astype(np.float64), P)
    assert relerr(Aa, (0.5 * P) @ V.astype(np.float64)) < 5e-6


@pytest.mark.parametrize("n,m,k,niter", [(1024, 256, 16, 20), (1000, 200, 10, 22), (640, 1204, 24, 12), (512, 384, 32, 12),
                                         (256, 128, 3, 25), (3108, 332, 7, 21), (128, 64, 1, 12)])
def test_tc_tiled_trace_parity_f32(ctx, n, m, k, niter):
    """Per-iteration W, H of the tcgen05 path against the Float64 oracle: <= 1e-4 relative (BASELINE north_star)."""
    X = synth.mixture(n, m, 3, seed=17, dtype=np.float32)
    W0, H0 = synth.philox_inits(23, 1, n, k, m, dtype=np.float32)
    Wt, Ht, ob = nb.trace(X, k, W0[0], H0[0], niter, ctx=ctx, engine=2)
    Wr, Hr, obr = oracle_trace(X, k, W0[0], H0[0], niter)
    for t in range(niter):
        assert relerr(Wt[t], Wr[t]) < RTOL32, ("W", t)
        assert relerr(Ht[t], Hr[t]) < RTOL32, ("H", t)
    assert np.allclose(ob, obr, rtol=2e-3, atol=1e-12 + 1e-5 * obr.max())


def test_tc2_random_shapes_match_scalar_pass(ctx):
    """Randomised ragged shapes (tools/tc2_stress.py) through the tcgen05 pass and the scalar-FMA pass of the tiled engine:
    own / reduction sizes that are not multiples of the 128 x 64 tile, 1..9 restarts (odd groups -> shadow units), k = 1..16."""
    rng = np.random.default_rng(123)
    for c in range(14):
        n = int(rng.choice([4, 8, 64, 128, 132, 260, 1000, 3108])) if c % 3 else 4 * int(rng.integers(1, 500))
        m = 4 * int(rng.integers(1, 400)) if c % 2 else int(rng.choice([4, 64, 68, 200, 332, 2000]))
        k, R, iters = int(rng.integers(1, 17)), int(rng.integers(1, 10)), int(rng.integers(1, 5))
        X = synth.mixture(n, m, min(k, 4), seed=c, dtype=np.float32)
        W0, H0 = synth.philox_inits(100 + c, R, n, k, m, dtype=np.float32)
        ctx.set_X(X)
        out = []
        for eng in (2, 4):
            b = ctx.batch(k, R)
            b.set_init(W0, H0)
            ctx.solve([b], nb.default_params(maxiter=iters, engine=eng))
            g = b.get()
            out.append((g["W"], g["H"]))
            b.close()
        assert relerr(out[0][0], out[1][0]) < 2e-5 and relerr(out[0][1], out[1][1]) < 2e-5, (c, n, m, k, R, iters)


def test_tc_generations_agree(ctx, tmp_path):
    """k <= 16 runs the second-generation tcgen05 pass (kl_tiled_tc2.cu); NMFK_TC_GEN=1 selects the first one (kl_tiled_tc.cu), which
    still serves 16 < k <= 32.  Same inputs through both (the first in a child process: the switch is read once per process), ragged
    sizes (edge tiles of the tensor-map TMA load, a partial last chunk, an odd number of restarts per group); the oracle comparisons
    of this file run the default, i.e. the second generation."""
    import os
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n, m, k, R, niter = 1156, 840, 12, 7, 6
    X = synth.mixture(n, m, 4, seed=31, dtype=np.float32)
    W0, H0 = synth.philox_inits(41, R, n, k, m, dtype=np.float32)
    np.savez(tmp_path / "in.npz", X=X, W0=W0, H0=H0)
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import nmfk_b200 as nb\n"
        "d = np.load(%r)\n"
        "with nb.Context(0) as ctx:\n"
        "    ctx.set_X(d['X'])\n"
        "    b = ctx.batch(%d, %d); b.set_init(d['W0'], d['H0'])\n"
        "    ctx.solve([b], nb.default_params(maxiter=%d, engine=2))\n"
        "    g = b.get(); np.savez(%r, W=g['W'], H=g['H'], iters=g['iters'])\n"
    ) % (os.path.join(ROOT, "nmfk.jl_b200", "python"), ROOT, str(tmp_path / "in.npz"), k, R, niter, str(tmp_path / "gen1.npz"))
    env = dict(os.environ, NMFK_TC_GEN="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    g1 = np.load(tmp_path / "gen1.npz")
    ctx.set_X(X)
    b = ctx.batch(k, R)
    b.set_init(W0, H0)
    ctx.solve([b], nb.default_params(maxiter=niter, engine=2))
    g2 = b.get()
    b.close()
    assert relerr(g2["W"], g1["W"]) < 2e-5 and relerr(g2["H"], g1["H"]) < 2e-5
    assert np.array_equal(g2["iters"], g1["iters"])


@pytest.mark.parametrize("n,m,k,R", [(2048, 512, 16, 9), (20000, 1000, 24, 5), (1000, 5000, 8, 6)])
def test_tc_tiled_restart_groups_match_scalar_pass(ctx, n, m, k, R):
    """Several restarts per CTA (ragged last group), sliced reductions and edge tiles: the tcgen05 pass and the
    scalar-FMA pass give the same factors after a fixed number of iterations."""
    X = synth.mixture(n, m, 4, seed=9, dtype=np.float32)
    ctx.set_X(X)
    res = {}
    for eng in (4, 2):
        b = ctx.batch(k, R)
        b.init_random(77)
        ctx.solve([b], nb.default_params(engine=eng, maxiter=8, normalize=0))
        res[eng] = b.get()
        b.close()
    assert relerr(res[2]["W"], res[4]["W"]) < 2e-5 and relerr(res[2]["H"], res[4]["H"]) < 2e-5
    assert np.allclose(res[2]["obj_norm"], res[4]["obj_norm"], rtol=1e-4)


def test_tc_tiled_frozen_restarts_and_stop_rule(ctx):
    """Restarts that stop early are frozen inside a restart group; iteration counts / stop reasons equal the
    scalar pass with the reference stop rule."""
    X = synth.mixture(512, 128, 3, seed=4, dtype=np.float32)
    ctx.set_X(X)
    res = {}
    for eng in (4, 2):
        b = ctx.batch(3, 7)
        b.init_random(5)
        ctx.solve([b], nb.default_params(engine=eng, maxiter=400))
        res[eng] = b.get()
        b.close()
    assert np.array_equal(res[2]["stop_reason"], res[4]["stop_reason"])
    assert np.max(np.abs(res[2]["iters"].astype(int) - res[4]["iters"].astype(int))) <= 10  # one check period
    assert np.allclose(res[2]["obj_norm"], res[4]["obj_norm"], rtol=5e-3)


# ---------------------------------------------------------------------------------------------
# Float64 tiled engine on the FP64 tensor pipe (kl_tiled_dmma.cu): engine=2 takes this path for Float64 data
# without NaN and 4 <= k <= 32; engine=4 forces the scalar-FMA pass.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m,k,R", [(20000, 600, 32, 3), (1001, 5001, 5, 4), (3000, 257, 13, 5), (130, 30000, 8, 2)])
def test_tiled_dmma_matches_scalar_pass(ctx, n, m, k, R):
    """Sliced reductions, odd sizes (unaligned X tiles), padded k, edge blocks: DMMA pass + DMMA objective against the
    scalar-FMA pass after a fixed number of iterations, and against the oracle for the first restart."""
    X = synth.mixture(n, m, 4, seed=21)
    W0, H0 = synth.philox_inits(91, R, n, k, m)
    ctx.set_X(X)
    res = {}
    for eng in (4, 2):
        b = ctx.batch(k, R)
        b.set_init(W0, H0)
        ctx.solve([b], nb.default_params(engine=eng, maxiter=12, normalize=0))
        res[eng] = b.get()
        b.close()
    assert relerr(res[2]["W"], res[4]["W"]) < 1e-10 and relerr(res[2]["H"], res[4]["H"]) < 1e-10
    assert np.allclose(res[2]["obj_norm"], res[4]["obj_norm"], rtol=1e-10)
    assert np.allclose(res[2]["obj_ssq"], res[4]["obj_ssq"], rtol=1e-10)
    if n * m <= 6e6:
        Xc = np.array(X, dtype=np.float64, order="F", copy=True)
        W, H, _ = o.nmf_multiplicative(Xc, k, Winit=W0[0].copy(), Hinit=H0[0].copy(), maxiter=12)
        assert relerr(res[2]["W"][0], W) < RTOL64 and relerr(res[2]["H"][0], H) < RTOL64


def test_execute_through_tensor_tiled_engines_same_decisions(ctx):
    """execute(X, ks, nNMF) with the tiled engine forced (engine=2: tcgen05 for Float32, DMMA for Float64) against the
    scalar-FMA tiled pass (engine=4) on a 3-source mixture: same kopt, same robustness ranking, fits equal to rounding."""
    for dt, tol in ((np.float32, 2e-3), (np.float64, 1e-6)):
        X = synth.mixture(1200, 400, 3, seed=31, dtype=dt)
        out = {}
        for eng in (4, 2):
            W, H, fit, rob, aic, kopt = nb.execute(X, range(2, 6), 8, seed=77, ctx=ctx, engine=eng, maxiter=400)
            out[eng] = (fit, rob, kopt)
        assert out[2][2] == out[4][2] == 3, (out[2][2], out[4][2])
        r2, r4 = np.asarray(out[2][1][1:5], dtype=float), np.asarray(out[4][1][1:5], dtype=float)
        assert np.array_equal(r2 > 0.5, r4 > 0.5)
        assert np.allclose(np.asarray(out[2][0][1:5], dtype=float), np.asarray(out[4][0][1:5], dtype=float), rtol=tol, atol=tol)


def test_cluster_means_and_variances_match_finalize(ctx):
    """Wmean / Hmean / Wvar / Hvar of finalize (NMFkFinalize.jl:68-74; what best=false returns) against the oracle, from the
    same restart solutions, sort order and labels."""
    for dt, tol in ((np.float64, 1e-12), (np.float32, 2e-6)):
        X = synth.mixture(90, 40, 3, seed=12, dtype=dt)
        k, R = 4, 7
        ctx.set_X(X)
        b = ctx.batch(k, R)
        b.init_random(321)
        ctx.solve([b], nb.default_params(maxiter=300))
        cl = b.cluster()
        sol = b.get()  # after cluster(): NaN entries zeroed like the reference does before finalize
        st = b.cluster_means(cl["order"], cl["labels"])
        b.close()
        Wa = [np.asarray(sol["W"][r], dtype=dt) for r in cl["order"]]
        Ha = [np.asarray(sol["H"][r], dtype=dt) for r in cl["order"]]
        Wm, Hm, _, Wv, Hv = o.finalize(Wa, Ha, np.asarray(cl["labels"]), False)
        for got, ref in ((st["W"], Wm), (st["H"], Hm), (st["Wvar"], Wv), (st["Hvar"], Hv)):
            assert got.shape == ref.shape
            assert np.allclose(got, ref, rtol=tol, atol=tol * max(1.0, float(np.max(np.abs(ref)))))


def test_execute_run_best_false_returns_cluster_means(ctx):
    """execute_run(...; best=false): Wa, Ha are the per-cluster means of finalize, phi / aic are derived from them
    (NMFkExecute.jl:637, 655-708) - against the oracle from the same initial factors."""
    X = synth.mixture(120, 30, 3, seed=3)
    k, R = 3, 6
    W0, H0 = synth.philox_inits(50, R, 120, k, 30)
    Wg, Hg, phig, robg, aicg = nb.execute_run(X, k, R, inits=(W0, H0), ctx=ctx, best=False)
    Wo, Ho, phio, robo, aico = o.execute_run(X.copy(), k, R, inits=[(W0[r].copy(), H0[r].copy()) for r in range(R)], best=False)
    assert relerr(Wg, Wo) < 1e-7 and relerr(Hg, Ho) < 1e-7
    assert abs(phig - phio) <= 1e-7 * max(phio, 1e-9) and abs(robg - robo) < 1e-8 and abs(aicg - aico) < 1e-5


def test_normalizevector_matches_oracle(ctx):
    """NMFmultiplicative(X, k; normalizevector) (NMFkMultiplicative.jl:27-31, 119-122) through the host mirror."""
    X = synth.mixture(60, 25, 3, seed=8)
    nv = np.random.default_rng(1).random(60) + 0.5
    W0, H0 = synth.philox_inits(4, 1, 60, 3, 25)
    Wg, Hg, og = nb.NMFmultiplicative(X, 3, Winit=W0[0], Hinit=H0[0], normalizevector=nv, maxiter=200, ctx=ctx)
    Wo, Ho, oo = o.nmf_multiplicative(X.copy(order="F"), 3, Winit=W0[0].copy(), Hinit=H0[0].copy(), normalizevector=nv, maxiter=200)
    assert relerr(Wg, Wo) < 1e-8 and relerr(Hg, Ho) < 1e-8 and abs(og - oo) <= 1e-7 * max(oo, 1e-12)
    with pytest.raises(nb.NMFkError):
        nb.NMFmultiplicative(X, 3, normalizevector=nv[:-1], ctx=ctx)
