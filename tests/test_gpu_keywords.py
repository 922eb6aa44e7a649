"""GPU parity of the keyword surface either side of the hot path (VERDICT r1 "missing" 5-9) - every case
against the CPU oracle from identical initial factors:
independent Winit / Hinit (the NMFkProgressive.jl:19 / NMFkMapping.jl:54 caller patterns), scalar / vector / matrix `weight`,
`normalizevector` through execute_run, the DArray method's own stop rule, clusterWmatrix (not forwarded to the restarts, best W
aliased by the centroids), acceptratio / acceptfactor / nanaction=:removed, best=false (nk = 1 and nk > 1), normalize=2."""
import numpy as np
import pytest

import nmfk_b200 as nb
from nmfk_b200 import synth
from oracle import nmfk_oracle as o

pytestmark = pytest.mark.gpu
STOP = {"maxiter": 1, "tol": 2, "reattempts": 3, "consistency": 4}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context()
    yield c
    c.close()


def _philox(seed, count):
    return np.random.Generator(np.random.Philox(key=seed)).random(count)


def test_hinit_only_with_hfixed_is_the_progressive_pattern(ctx):
    """NMFkProgressive.jl:19: execute(X, nk, nNMF; Hinit=H, Hfixed=true): W is drawn per restart (the FIRST numbers of the
    restart's stream, NMFkMultiplicative.jl:37-45), H is shared and never changes; no signal reordering (:305-307)."""
    n, m, k, R = 80, 30, 3, 5
    X = synth.mixture(n, m, 3, seed=6)
    Hfix = np.asfortranarray(np.random.default_rng(2).random((k, m)))
    inits = [(_philox(300 + i, n * k).reshape((n, k), order="F"), Hfix.copy()) for i in range(1, R + 1)]
    Wg, Hg, fg, rg, ag = nb.execute_k(X, k, R, seed=300, Hinit=Hfix, Hfixed=True, ctx=ctx)
    Wo, Ho, fo, ro, ao = o.execute_k(X.copy(), k, R, inits=inits, Hfixed=True)
    assert np.array_equal(Hg, Hfix)
    assert relerr(Wg, Wo) < 1e-7 and abs(fg - fo) <= 1e-7 * fo and abs(rg - ro) < 1e-8 and abs(ag - ao) < 1e-5
    # ... and through the k sweep entry (single k)
    W, H, fit, rob, aic, kopt = nb.execute(X, [k], R, seed=300, Hinit=Hfix, Hfixed=True, ctx=ctx)
    assert relerr(W[k], Wo) < 1e-7 and np.array_equal(H[k], Hfix)


def test_winit_only_with_wfixed_is_the_mapping_pattern(ctx):
    """NMFkMapping.jl:54: Winit + Wfixed: H is drawn from the start of the restart's stream (W consumes nothing)."""
    n, m, k, R = 60, 40, 4, 4
    X = synth.mixture(n, m, 4, seed=16)
    Wfix = np.asfortranarray(np.random.default_rng(5).random((n, k)))
    inits = [(Wfix.copy(), _philox(900 + i, k * m).reshape((k, m), order="F")) for i in range(1, R + 1)]
    Wg, Hg, fg, rg, ag = nb.execute_run(X, k, R, seed=900, Winit=Wfix, Wfixed=True, ctx=ctx)
    Wo, Ho, fo, ro, ao = o.execute_run(X.copy(), k, R, inits=inits, Wfixed=True)
    assert np.array_equal(Wg, Wfix)
    assert relerr(Hg, Ho) < 1e-7 and abs(fg - fo) <= 1e-7 * fo and abs(rg - ro) < 1e-8
    # NMFmultiplicative directly with one factor given
    W1, H1, o1 = nb.NMFmultiplicative(X, k, Winit=Wfix, seed=77, maxiter=50, ctx=ctx)
    W2, H2, o2 = o.nmf_multiplicative(X.copy(), k, Winit=Wfix.copy(), Hinit=_philox(77, k * m).reshape((k, m), order="F"), maxiter=50)
    assert relerr(W1, W2) < 1e-9 and relerr(H1, H2) < 1e-9 and abs(o1 - o2) <= 1e-8 * o2


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["scalar", "rows", "cols", "matrix"])
def test_weight_in_the_objective(ctx, kind, dt):
    """weight (NMFkExecute.jl:484, NMFkMultiplicative.jl:74,125): scalar != 1, a vector of length n (rows), 1 x m (columns),
    n x m - it only enters the objective, i.e. the stop machine and the reported sum of squares."""
    n, m, k = 96, 40, 3
    X = synth.mixture(n, m, 3, seed=14, dtype=dt)
    rng = np.random.default_rng(4)
    w = {"scalar": 3.5, "rows": rng.random(n) + 0.5, "cols": rng.random((1, m)) + 0.5, "matrix": rng.random((n, m)) + 0.5}[kind]
    wd = w if np.isscalar(w) else np.asarray(w, dtype=dt)
    W0, H0 = synth.philox_inits(19, 1, n, k, m, dtype=dt)
    inf = {}
    Wo, Ho, oo = o.nmf_multiplicative(np.asfortranarray(X.astype(np.float64)), k, Winit=W0[0].astype(np.float64),
                                      Hinit=H0[0].astype(np.float64), weight=wd if np.isscalar(wd) else wd.astype(np.float64),
                                      maxiter=3000, info=inf)
    ctx.set_X(X)
    b = ctx.batch(k, 1)
    b.set_init(W0, H0)
    from nmfk_b200.api import _params_from_kw
    ctx.solve([b], _params_from_kw({"weight": wd, "maxiter": 3000}, ctx, normalize=0))
    g = b.get()
    b.close()
    ctx.set_weight(None)
    tol = 1e-7 if dt == np.float64 else 2e-3
    if dt == np.float64:
        assert g["iters"][0] == inf["iters"] and g["stop_reason"][0] == STOP[inf["stop_reason"]]
        assert relerr(g["W"][0], Wo) < 1e-7 and relerr(g["H"][0], Ho) < 1e-7
    else:
        assert abs(int(g["iters"][0]) - inf["iters"]) <= 20
    assert abs(g["obj_ssq"][0] - oo) <= tol * oo + 1e-12, (g["obj_ssq"][0], oo)


def test_normalizevector_through_execute_run(ctx):
    """NMFkProgressive.jl:82-96 passes normalizevector down execute -> execute_run -> NMFmultiplicative: rows of X divided for
    the solve (:27-31), W scaled back and the objective taken against X (:119-125), then the H-row normalisation."""
    n, m, k, R = 60, 25, 3, 4
    X = synth.mixture(n, m, 3, seed=8)
    nv = np.random.default_rng(1).random(n) + 0.5
    W0, H0 = synth.philox_inits(4, R, n, k, m)
    Wg, Hg, fg, rg, ag = nb.execute_run(X, k, R, inits=(W0, H0), normalizevector=nv, maxiter=300, ctx=ctx)
    Wo, Ho, fo, ro, ao = o.execute_run(X.copy(order="F"), k, R, inits=[(W0[r].copy(), H0[r].copy()) for r in range(R)],
                                       normalizevector=nv, maxiter=300)
    assert relerr(Wg, Wo) < 1e-7 and relerr(Hg, Ho) < 1e-7 and abs(fg - fo) <= 1e-7 * fo and abs(rg - ro) < 1e-8
    Wd, Hd, od = nb.NMFmultiplicative(X, k, Winit=W0[0], Hinit=H0[0], normalizevector=nv, weight=2.0, maxiter=200, ctx=ctx)
    We, He, oe = o.nmf_multiplicative(X.copy(order="F"), k, Winit=W0[0].copy(), Hinit=H0[0].copy(), normalizevector=nv, weight=2.0,
                                      maxiter=200)
    assert relerr(Wd, We) < 1e-8 and relerr(Hd, He) < 1e-8 and abs(od - oe) <= 1e-7 * max(oe, 1e-12)
    with pytest.raises(nb.NMFkError):
        nb.NMFmultiplicative(X, k, normalizevector=nv[:-1], ctx=ctx)


@pytest.mark.parametrize("engine", [0, 2])
@pytest.mark.parametrize("case", ["readme", "noisy"])
def test_darray_stop_rule(ctx, engine, case):
    """NMFmultiplicative(::DArray) (NMFkMultiplicative.jl:129-197): no tolOF / baditers / reattempts - where the dense method
    gives up on `reattempts` the distributed one goes on until objvalue < tol (README problem: iteration 1040) or maxiter (noisy
    mixture); identical iteration count, stop reason, factors and objective vs the oracle."""
    k = 3
    if case == "readme":
        X = synth.readme_bss()
    else:
        X = synth.mixture(40, 12, 3, seed=2) + 0.05 * np.random.default_rng(0).random((40, 12))
    n, m = X.shape
    W0, H0 = synth.philox_inits(100, 1, n, k, m)
    inf_dense, inf = {}, {}
    o.nmf_multiplicative(X.copy(), k, Winit=W0[0].copy(), Hinit=H0[0].copy(), maxiter=2000, info=inf_dense)
    Wo, Ho, oo = o.nmf_multiplicative_darray(X.copy(), k, Winit=W0[0].copy(), Hinit=H0[0].copy(), maxiter=2000, info=inf)
    assert inf_dense["stop_reason"] == "reattempts" and inf_dense["iters"] < inf["iters"]
    assert (inf["iters"], inf["stop_reason"]) == ((1040, "tol") if case == "readme" else (2000, "maxiter"))
    ctx.set_X(X)
    b = ctx.batch(k, 1)
    b.set_init(W0, H0)
    ctx.solve([b], nb.default_params(maxiter=2000, stop_rule=1, stopconv=10000, normalize=0, engine=engine))
    g = b.get()
    b.close()
    assert g["iters"][0] == inf["iters"] and g["stop_reason"][0] == STOP[inf["stop_reason"]]
    assert relerr(g["W"][0], Wo) < 1e-7 and relerr(g["H"][0], Ho) < 1e-7 and abs(g["obj_ssq"][0] - oo) <= 1e-6 * oo + 1e-18
    Wd, Hd, od = nb.NMFmultiplicative_darray(X, k, Winit=W0[0], Hinit=H0[0], maxiter=2000, ctx=ctx, engine=engine)
    assert relerr(Wd, Wo) < 1e-7


def test_cluster_wmatrix_is_not_forwarded_and_best_w_is_aliased(ctx):
    """execute_run(...; clusterWmatrix=true): the restarts are still normalised by the rows of H (the keyword is consumed by
    execute_run, NMFkExecute.jl:483, 516-540), the columns of W are clustered (:620-621), and - reference behaviour -
    clustersolutions accumulates its centroids IN the best solution's W (NMFkCluster.jl:453-455), which Wbest then reads."""
    n, m, k, R = 70, 26, 3, 6
    X = synth.mixture(n, m, 3, seed=3)
    W0, H0 = synth.philox_inits(50, R, n, k, m)
    dg, do = {}, {}
    Wg, Hg, fg, rg, ag = nb.execute_run(X, k, R, inits=(W0, H0), clusterWmatrix=True, ctx=ctx, details=dg)
    Wo, Ho, fo, ro, ao = o.execute_run(X.copy(), k, R, inits=[(W0[r].copy(), H0[r].copy()) for r in range(R)], clusterWmatrix=True,
                                       details=do)
    assert np.array_equal(dg["labels"], do["labels"]) and np.array_equal(dg["order"], do["idxsort"])
    assert np.allclose(Hg.sum(axis=1), 1.0, atol=1e-12)  # rows of H sum to one although clusterWmatrix = true
    assert relerr(Wg, Wo) < 1e-7 and relerr(Hg, Ho) < 1e-7
    assert abs(fg - fo) <= 1e-7 * fo and abs(rg - ro) < 1e-8 and abs(ag - ao) <= 1e-6 * abs(ao)
    assert np.allclose(dg["clustersil"], do["clustersil"][:, 0], atol=1e-9)
    assert relerr(dg["Wmean"], do["Wmean"]) < 1e-9 and relerr(dg["Hmean"], do["Hmean"]) < 1e-9
    # the one-call form takes the same route
    W1, H1, f1, r1, a1 = nb.execute_run(X, k, R, inits=(W0, H0), clusterWmatrix=True, ctx=ctx)
    assert relerr(W1, Wo) < 1e-7 and abs(r1 - ro) < 1e-8
    # execute_singlerun_compute called directly DOES honour its own clusterWmatrix keyword (:796-799): normalize = 2
    Ws, Hs, os_ = nb.execute_singlerun(X, k, Winit=W0[0], Hinit=H0[0], clusterWmatrix=True, ctx=ctx)
    We, He, oe = o.execute_singlerun_compute(X.copy(), k, Winit=W0[0].copy(), Hinit=H0[0].copy(), clusterWmatrix=True)
    assert np.allclose(Ws.sum(axis=0), 1.0, atol=1e-12) and relerr(Ws, We) < 1e-7 and relerr(Hs, He) < 1e-7


@pytest.mark.parametrize("opts", [dict(acceptratio=0.5), dict(acceptfactor=1.5), dict(acceptratio=0.7, acceptfactor=3.0),
                                  dict(best=False), dict(acceptratio=0.6, best=False)])
def test_solution_filters_and_best_false(ctx, opts):
    """acceptratio / acceptfactor (NMFkExecute.jl:551-565) and best=false (:655-658) against the oracle."""
    n, m, k, R = 90, 30, 3, 8
    X = synth.mixture(n, m, 3, seed=12) + 0.05 * np.random.default_rng(0).random((n, m))
    W0, H0 = synth.philox_inits(61, R, n, k, m)
    dg, do = {}, {}
    Wg, Hg, fg, rg, ag = nb.execute_run(X, k, R, inits=(W0, H0), ctx=ctx, details=dg, maxiter=500, **opts)
    Wo, Ho, fo, ro, ao = o.execute_run(X.copy(), k, R, inits=[(W0[r].copy(), H0[r].copy()) for r in range(R)], details=do,
                                       maxiter=500, **opts)
    assert np.array_equal(dg["order"], do["idxsort"][do["idxsol"]])
    assert np.array_equal(dg["labels"], do["labels"])
    assert relerr(Wg, Wo) < 1e-7 and relerr(Hg, Ho) < 1e-7 and abs(fg - fo) <= 1e-7 * fo and abs(rg - ro) < 1e-8


def test_nanaction_removed_and_best_false_with_one_signal(ctx):
    """nanaction=:removed drops solutions that hold NaN (:581-596) - position-wise, like the reference; nk = 1 with best=false
    returns the first kept restart (finalize(WBig[idxsol], HBig[idxsol]), :646-650)."""
    n, m, R = 40, 12, 5
    X = synth.mixture(n, m, 2, seed=5)
    ctx.set_X(X)
    # solutions imported as they are: restart 1 holds a NaN
    rng = np.random.default_rng(9)
    k = 2
    W = rng.random((R, n, k))
    H = rng.random((R, k, m))
    H[1, 0, 3] = np.nan
    obj = np.array([3.0, 1.0, 2.0, 5.0, 4.0])
    b = ctx.import_solutions(H, obj, W=W)
    kept = b.select(nanaction="removed")
    # sorted order is [1, 2, 0, 4, 3]; idxnan[1] = false masks POSITION 1 (restart 2), as the reference's mixed indexing does
    assert list(kept) == [1, 0, 4, 3]
    cl = b.cluster()
    assert cl["labels"].shape == (k, 4)
    b.close()
    # nk = 1
    W0, H0 = synth.philox_inits(7, R, n, 1, m)
    Wg, Hg, fg, rg, ag = nb.execute_run(X, 1, R, inits=(W0, H0), ctx=ctx, best=False, maxiter=100)
    Wo, Ho, fo, ro, ao = o.execute_run(X.copy(), 1, R, inits=[(W0[r].copy(), H0[r].copy()) for r in range(R)], best=False, maxiter=100)
    assert rg == ro == 1 and relerr(Wg, Wo) < 1e-8 and relerr(Hg, Ho) < 1e-8 and abs(fg - fo) <= 1e-8 * fo
