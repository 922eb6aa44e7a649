"""NMFsparsity (NMFkSparsity.jl:1-113; method=:sparsity) on the GPU against the oracle's line-by-line restatement:
per-iteration objective, final factors, the relative-change stop rule, through NMFsparsity and through execute_run."""
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


@pytest.mark.parametrize("cf,sparsity,n,m,k,dt", [("ed", 1.0, 120, 50, 3, np.float64), ("ed", 0.0, 300, 64, 10, np.float64),
                                                  ("kl", 0.1, 90, 40, 4, np.float64), ("is", 0.01, 64, 32, 2, np.float64),
                                                  ("ed", 0.5, 256, 128, 16, np.float32), ("kl", 0.0, 200, 96, 20, np.float64)])
def test_nmfsparsity_matches_oracle(ctx, cf, sparsity, n, m, k, dt):
    X = synth.mixture(n, m, 3, seed=21, dtype=dt) + dt(0.01)
    W0, H0 = synth.philox_inits(5, 1, n, k, m, dtype=dt)
    niter = 60
    W, H, obj = nb.NMFsparsity(X, k, Winit=W0[0], Hinit=H0[0], cost_function=cf, sparsity=sparsity, maxiter=niter, ctx=ctx)
    inf = {}
    Wo, Ho, objo = o.nmf_sparsity(X.astype(np.float64), k, Winit=W0[0].astype(np.float64), Hinit=H0[0].astype(np.float64), cost_function=cf,
                                  sparsity=sparsity, maxiter=niter, info=inf)
    tol = 1e-8 if dt == np.float64 else 2e-4
    assert relerr(W, Wo) < tol and relerr(H, Ho) < tol, (relerr(W, Wo), relerr(H, Ho))
    assert abs(obj - objo) <= (1e-7 if dt == np.float64 else 5e-3) * objo + 1e-12
    assert np.allclose(np.linalg.norm(W.astype(np.float64), axis=0), 1.0, atol=1e-6)  # unit-norm columns of W (:86)


def test_fractional_beta_stop_rule_and_execute_run(ctx):
    X = synth.mixture(80, 30, 3, seed=2) + 0.05
    W0, H0 = synth.philox_inits(9, 4, 80, 3, 30)
    # fractional beta (:65-67, :80-82, :96-97)
    W, H, obj = nb.NMFsparsity(X, 3, Winit=W0[0], Hinit=H0[0], beta_divergence=1.5, sparsity=0.2, maxiter=40, ctx=ctx)
    Wo, Ho, objo = o.nmf_sparsity(X.copy(), 3, Winit=W0[0].copy(), Hinit=H0[0].copy(), beta_divergence=1.5, sparsity=0.2, maxiter=40)
    assert relerr(W, Wo) < 1e-7 and relerr(H, Ho) < 1e-7
    # relative-change stop (:101-106) with a loose tol: same iteration count
    ctx.set_X(X)
    from nmfk_b200.api import _params_from_kw
    p = _params_from_kw(dict(method="sparsity", sparsity=0.3), ctx, maxiter=5000, tol=1e-6, normalize=0)
    b = ctx.batch(3, 4)
    b.set_init(W0, H0)
    ctx.solve([b], p)
    g = b.get()
    b.close()
    for r in range(4):
        inf = {}
        Wr, Hr, _ = o.nmf_sparsity(X.copy(), 3, Winit=W0[r].copy(), Hinit=H0[r].copy(), sparsity=0.3, maxiter=5000, tol=1e-6, info=inf)
        assert g["iters"][r] == inf["iters"] and g["stop_reason"][r] == 2, (r, g["iters"][r], inf)
        assert relerr(g["W"][r], Wr) < 1e-7 and relerr(g["H"][r], Hr) < 1e-7
    # through execute_run with the reference's keyword: objective = normnan(X - W*H), rows of H sum to one
    Wg, Hg, fg, rg, ag = nb.execute_run(X, 3, 4, inits=(W0, H0), method="sparsity", sparsity=0.3, maxiter=200, ctx=ctx)
    assert np.allclose(Hg.sum(axis=1), 1.0, atol=1e-10) and np.linalg.norm(X - Wg @ Hg) == pytest.approx(float(fg), rel=1e-8)
