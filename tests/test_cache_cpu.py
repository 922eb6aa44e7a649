"""Result-cache contract (NMFkExecute.jl:264-303, 323-327, 650-654; NMFkIO.jl:45-128) on the CPU: file names,
key names, hit / miss / inconsistent / loadonly logic, fit re-derivation, old-name fallback and renaming.  The solver is a stub
(`runner`): nothing here needs a GPU."""
import os

import numpy as np
import pytest

from nmfk_b200 import cache
from oracle import nmfk_oracle as o


def _problem():
    rng = np.random.default_rng(0)
    W, H = rng.random((12, 3)), rng.random((3, 7))
    return W @ H, W, H


def _runner_factory(W, H, calls):
    def runner(X, nk, nNMF, **kw):
        calls.append((nk, nNMF, dict(kw)))
        return W.copy(), H.copy(), cache.normnan(X - W @ H), 0.9, -12.5
    return runner


def test_names_and_keys(tmp_path):
    assert os.path.basename(cache.result_filename(".", "case", 12, 7, 3, 10, ".jld")) == "case_12_7_3_10.jld"
    assert os.path.basename(cache.old_result_filename(".", "case", 3, 10, ".jld")) == "case-3-10.jld"
    assert os.path.basename(cache.all_filename(".", "case", 12, 7, 3, 10, ".jld")) == "case_12_7_3_10-all.jld"
    assert cache.RESULT_KEYS == ("W", "H", "fit", "robustness", "aic")
    assert set(cache.ALL_KEYS) >= {"Wmean", "Hvar", "Wbest", "Cluster Silhouettes", "Cluster assignments", "Cluster centroids"}


def test_execute_cached_miss_then_hit(tmp_path):
    X, W, H = _problem()
    calls = []
    run = _runner_factory(W, H, calls)
    kw = dict(runner=run, signalorder=o.signalorder, resultdir=str(tmp_path), casefilename="t")
    W1, H1, f1, r1, a1 = cache.execute_k_cached(X, 3, 10, **kw)
    assert len(calls) == 1 and os.path.isfile(tmp_path / "t_12_7_3_10.npz")
    so = o.signalorder(W, H)
    assert np.array_equal(W1, W[:, so]) and np.array_equal(H1, H[so, :])
    W2, H2, f2, r2, a2 = cache.execute_k_cached(X, 3, 10, **kw)  # hit: the solver is not called again
    assert len(calls) == 1 and np.array_equal(W2, W1) and (f2, r2, a2) == (f1, r1, a1)
    # another nNMF is another file
    cache.execute_k_cached(X, 3, 20, **kw)
    assert len(calls) == 2 and os.path.isfile(tmp_path / "t_12_7_3_20.npz")
    # load=false always runs, save=false never writes
    cache.execute_k_cached(X, 3, 30, load=False, save=False, **kw)
    assert len(calls) == 3 and not os.path.isfile(tmp_path / "t_12_7_3_30.npz")


def test_inconsistent_fit_is_rederived_and_saved(tmp_path):
    X, W, H = _problem()
    calls = []
    kw = dict(runner=_runner_factory(W, H, calls), signalorder=o.signalorder, resultdir=str(tmp_path), casefilename="t")
    cache._write(cache.result_filename(str(tmp_path), "t", 12, 7, 3, 10), W=W, H=H, fit=123.0, robustness=0.5, aic=1.0)
    Wl, Hl, fit, rob, aic = cache.execute_k_cached(X, 3, 10, **kw)
    assert len(calls) == 0 and abs(fit - cache.normnan(X - W @ H)) < 1e-12 and rob == 0.5
    assert abs(cache._read(cache.result_filename(str(tmp_path), "t", 12, 7, 3, 10), ("fit",))[0] - fit) < 1e-12  # saved again (:278-281)


def test_inconsistent_shapes_rerun_and_loadonly(tmp_path):
    X, W, H = _problem()
    calls = []
    kw = dict(runner=_runner_factory(W, H, calls), signalorder=o.signalorder, resultdir=str(tmp_path), casefilename="t")
    cache._write(cache.result_filename(str(tmp_path), "t", 12, 7, 3, 10), W=W[:, :2], H=H[:2], fit=1.0, robustness=0.5, aic=1.0)
    cache.execute_k_cached(X, 3, 10, **kw)
    assert len(calls) == 1  # "contains inconsistent results; runs will be executed" (:287-291)
    Wl, Hl, fit, rob, aic = cache.execute_k_cached(X, 4, 10, loadonly=True, **kw)  # missing + loadonly: no run, empty result
    assert len(calls) == 1 and Wl.shape == (0, 0) and fit == np.inf and rob == -1 and aic == -np.inf
    assert not os.path.isfile(tmp_path / "t_12_7_4_10.npz")
    with pytest.raises(ValueError):
        cache.execute_k_cached(np.empty((0, 3)), 2, 10, **kw)


def test_old_name_fallback_load_and_rename(tmp_path):
    X, W, H = _problem()
    cache._write(cache.old_result_filename(str(tmp_path), "t", 3, 10), W=W, H=H, fit=cache.normnan(X - W @ H), robustness=0.7, aic=2.0)
    calls = []
    out = cache.execute_k_cached(X, 3, 10, runner=_runner_factory(W, H, calls), signalorder=o.signalorder, resultdir=str(tmp_path),
                                 casefilename="t")
    assert len(calls) == 0 and out[3] == 0.7  # found under the old convention (:266-269)
    os.remove(tmp_path / "t_12_7_3_10.npz") if os.path.isfile(tmp_path / "t_12_7_3_10.npz") else None
    Wl, Hl, fit, rob, aic = cache.load(3, 10, resultdir=str(tmp_path), casefilename="t", signalorder=o.signalorder)
    assert rob == 0.7 and os.path.isfile(tmp_path / "t_12_7_3_10.npz") and not os.path.isfile(tmp_path / "t-3-10.npz")  # renamed (:92-95)
    Wm, Hm, fm, rm, am = cache.load(5, 10, resultdir=str(tmp_path), casefilename="t")
    assert Wm.shape == (0, 0) and np.isnan(fm)
    W_, H_, f_, r_, a_, kopt = cache.load_range([3], 10, resultdir=str(tmp_path), casefilename="t", signalorder=o.signalorder, getk=o.getk)
    assert kopt == 3
    assert cache.save(W, H, 1.0, 0.7, 2.0, 3, 10, resultdir=str(tmp_path), casefilename="t") is False  # never overwrites (:121-123)


def test_all_file_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    R, n, k, m = 4, 6, 2, 5
    det = dict(W=rng.random((R, n, k)), H=rng.random((R, k, m)), Wmean=rng.random((n, k)), Hmean=rng.random((k, m)), Wvar=rng.random((n, k)),
               Hvar=rng.random((k, m)), Wbest=rng.random((n, k)), Hbest=rng.random((k, m)), fit=rng.random(R), clustersil=rng.random(k),
               labels=np.tile(np.arange(1, k + 1)[:, None], (1, R)), centroids=rng.random((k, m)))
    fn = cache.save_all(det, (n, m), k, R, resultdir=str(tmp_path), casefilename="c")
    assert os.path.basename(fn) == "c_6_5_2_4-all.npz"
    with np.load(fn) as f:
        assert set(f.files) == set(cache.ALL_KEYS) and f["Cluster Silhouettes"].shape == (k, 1)
    WBig, HBig, obj = cache.load_all((n, m), k, R, resultdir=str(tmp_path), casefilename="c")
    assert np.array_equal(WBig, det["W"]) and np.array_equal(obj, det["fit"])
    assert cache.load_all((n, m), 3, R, resultdir=str(tmp_path), casefilename="c") is None
