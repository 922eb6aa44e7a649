"""Result-cache contract of the hot path: the on-disk step either side of `execute`.

The reference stores one file per (k, nNMF) and optionally one "-all" file with every restart
(/root/reference/src/NMFkExecute.jl:264-303, 323-327, 492-506, 650-654; /root/reference/src/NMFkIO.jl:45-128):

    <casefilename>_<n>_<m>_<k>_<nNMF>.jld        keys "W", "H", "fit", "robustness", "aic"          (W[:,so], H[so,:])
    <casefilename>_<n>_<m>_<k>_<nNMF>-all.jld    keys "W", "H", "Wmean", "Hmean", "Wvar", "Hvar", "Wbest", "Hbest", "fit",
                                                      "Cluster Silhouettes", "Cluster assignments", "Cluster centroids"
    <casefilename>-<k>-<nNMF>.jld                the old naming convention, still looked up (and renamed by `load`)

JLD is Julia's HDF5 dialect; neither Julia nor an HDF5 library exists in this environment, so THIS mirror keeps the contract
(file names, key names, shapes, the hit / miss / inconsistent-shape / fit re-derivation logic, `loadonly`, the old-name
fallback, `ordersignals`) on a NumPy `.npz` container - the Julia shim (julia/NMFkB200) applies the same logic with `JLD.save`
/ `JLD.load`, so result directories written by the reference load there unchanged.  The file extension is the only difference
and is a parameter (`ext`).  Nothing here touches the device: the solver is passed in (`runner`), which is how the CPU test
suite exercises it."""
from __future__ import annotations

import os
import re
from typing import Callable, Optional

import numpy as np

RESULT_KEYS = ("W", "H", "fit", "robustness", "aic")
ALL_KEYS = ("W", "H", "Wmean", "Hmean", "Wvar", "Hvar", "Wbest", "Hbest", "fit", "Cluster Silhouettes", "Cluster assignments",
            "Cluster centroids")
EPS_F16 = float(np.finfo(np.float16).eps)  # `abs(fit - fitquality) > eps(Float16)` (NMFkExecute.jl:277)


def result_filename(resultdir, casefilename, n, m, k, nNMF, ext=".npz"):
    """NMFkExecute.jl:265, 324: "$(casefilename)_$(size(X,1))_$(size(X,2))_$(nk)_$(nNMF).jld"."""
    return os.path.join(resultdir, "%s_%d_%d_%d_%d%s" % (casefilename, n, m, k, nNMF, ext))


def old_result_filename(resultdir, casefilename, k, nNMF, ext=".npz"):
    """NMFkExecute.jl:268: "$(casefilename)-$(nk)-$(nNMF).jld" (old convention)."""
    return os.path.join(resultdir, "%s-%d-%d%s" % (casefilename, k, nNMF, ext))


def all_filename(resultdir, casefilename, n, m, k, nNMF, ext=".npz"):
    """NMFkExecute.jl:494, 651: "...-all.jld"."""
    return os.path.join(resultdir, "%s_%d_%d_%d_%d-all%s" % (casefilename, n, m, k, nNMF, ext))


def _scalar(v):
    return v.item() if isinstance(v, np.ndarray) and v.ndim == 0 else v


def _read(filename, keys):
    with np.load(filename, allow_pickle=False) as f:
        return tuple(_scalar(f[k]) for k in keys)


def _write(filename, **items):
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)  # joinpathcheck / recursivemkdir
    with open(filename, "wb") as fh:
        np.savez(fh, **items)


def save(W, H, fitquality, robustness, aic, nk: int, nNMF: int = 10, *, resultdir=".", casefilename="nmfk", filename="", ext=".npz"):
    """`NMFk.save(W, H, fitquality, robustness, aic, nk, nNMF; ...)` NMFkIO.jl:112-124: never overwrites."""
    if casefilename != "" and filename == "":
        filename = result_filename(resultdir, casefilename, W.shape[0], H.shape[1], nk, nNMF, ext)
    if os.path.isfile(filename):
        return False  # "File named ... already exists!"
    _write(filename, W=W, H=H, fit=fitquality, robustness=robustness, aic=aic)
    return True


def load(nk: int, nNMF: int = 10, *, dtype=np.float64, resultdir=".", casefilename="nmfk", filename="", ordersignals=True, ext=".npz",
         signalorder: Optional[Callable] = None):
    """`NMFk.load(nk, nNMF; ...)` NMFkIO.jl:45-102: looks for the old name, then "<case>_<nk>_<nNMF>", then the size-encoded
    "<case>_<n>_<m>_<nk>_<nNMF>"; reorders the signals; renames a file found under the old convention.  A missing file gives
    empty matrices and NaN scores."""
    if casefilename != "" and filename == "":
        filename = old_result_filename(resultdir, casefilename, nk, nNMF, ext)
        if not os.path.isfile(filename):
            filename = os.path.join(resultdir, "%s_%d_%d%s" % (casefilename, nk, nNMF, ext))
            if not os.path.isfile(filename):
                pat = re.compile(re.escape(casefilename) + r"_(\d+)_(\d+)_%d_%d%s$" % (nk, nNMF, re.escape(ext)))
                try:
                    for f in sorted(os.listdir(resultdir)):
                        if pat.match(f):
                            filename = os.path.join(resultdir, f)
                            break
                except OSError:
                    pass
    if os.path.isfile(filename):
        W, H, fit, rob, aic = _read(filename, RESULT_KEYS)
        so = np.arange(W.shape[1])
        if ordersignals:
            so = signalorder(W, H) if signalorder is not None else np.argsort(-np.array([np.sum(np.outer(W[:, i], H[i, :])) for i in
                                                                                         range(W.shape[1])]), kind="stable")
        if filename == old_result_filename(resultdir, casefilename, nk, nNMF, ext):
            os.replace(filename, result_filename(resultdir, casefilename, W.shape[0], H.shape[1], nk, nNMF, ext))  # :92-95
        return W[:, so], H[so, :], fit, rob, aic
    return np.empty((0, 0), dtype=dtype), np.empty((0, 0), dtype=dtype), np.nan, np.nan, np.nan


def load_range(nkrange, nNMF: int = 10, *, cutoff=0.5, strict=True, getk: Optional[Callable] = None, **kw):
    """`NMFk.load(nkrange, nNMF; cutoff)` NMFkIO.jl:4-38 -> (W, H, fitquality, robustness, aic, kopt) keyed by k."""
    W, H, fit, rob, aic = {}, {}, {}, {}, {}
    for k in nkrange:
        W[k], H[k], fit[k], rob[k], aic[k] = load(k, nNMF, **kw)
    robs = [rob[k] for k in nkrange]
    if getk is None:
        from .api import getk as _getk
        getk = _getk
    return W, H, fit, rob, aic, getk(list(nkrange), robs, cutoff, strict)


def normnan(A):
    v = np.asarray(A)[~np.isnan(A)]
    return float(np.sqrt(np.sum(v.astype(np.float64) ** 2)))


def execute_k_cached(X, nk: int, nNMF: int = 10, *, runner: Callable, signalorder: Callable, resultdir=".", casefilename="nmfk",
                     loadonly=False, load=True, save=True, ordersignals=True, ext=".npz", **kw):
    """`NMFk.execute(X, nk::Integer, nNMF; resultdir, casefilename, loadonly, load, save, ordersignals, ...)`
    NMFkExecute.jl:236-329 WITH the file cache.  runner(X, nk, nNMF, **kw) -> (W, H, fitquality, robustness, aic) stands for
    execute_run (nmfk_b200.execute_run on the GPU); signalorder(W, H) -> 0-based order.
    Returns (W[:,so], H[so,:], fitquality, robustness, aic)."""
    X = np.asarray(X)
    if X.size == 0:
        raise ValueError("Input array has a zero dimension! Array size=%s" % (X.shape,))  # :242-244
    runflag = True
    if loadonly:  # :245-251
        load, save, runflag = True, False, False
    n, m = X.shape
    execute_ordersignals = True
    W = H = None
    fitquality = robustness = aic = None
    if load:  # :264-303
        filename = result_filename(resultdir, casefilename, n, m, nk, nNMF, ext)
        if not os.path.isfile(filename):
            filename = old_result_filename(resultdir, casefilename, nk, nNMF, ext)
        if os.path.isfile(filename):
            W, H, fitquality, robustness, aic = _read(filename, RESULT_KEYS)
            if W.shape == (n, nk) and H.shape == (nk, m):
                fit = normnan(X - W @ H)
                if abs(fit - fitquality) > EPS_F16:  # "Fit quality is not consistent": keep the new fit and save it again
                    fitquality = fit
                    save = True
                else:
                    save = False
                runflag = False
            else:  # inconsistent results: runs will be executed
                W = H = None
        elif loadonly:
            W, H = np.empty((0, 0), dtype=X.dtype), np.empty((0, 0), dtype=X.dtype)
            fitquality, robustness, aic = np.inf, -1, -np.inf
            execute_ordersignals = False
    if "Wfixed" in kw or "Hfixed" in kw:  # :305-307
        ordersignals = False
    if runflag and not (loadonly and W is not None):
        W, H, fitquality, robustness, aic = runner(X, nk, nNMF, **kw)  # :309
    if execute_ordersignals:
        so = np.asarray(signalorder(W, H)) if ordersignals else np.arange(W.shape[1])  # :311-318
    else:
        so = np.arange(0)
    if save:  # :323-327 (JLD.save overwrites)
        _write(result_filename(resultdir, casefilename, n, m, nk, nNMF, ext), W=W[:, so], H=H[so, :], fit=fitquality,
               robustness=robustness, aic=aic)
    return W[:, so], H[so, :], fitquality, robustness, aic


def save_all(details: dict, X_shape, nk: int, nNMF: int, *, resultdir=".", casefilename="nmfk", ext=".npz"):
    """The "-all" file of execute_run(...; saveall=true) NMFkExecute.jl:650-654 from the `details` dict that
    nmfk_b200.execute_run fills: every restart's W and H, the per-cluster means / variances of finalize, the best restart,
    the objective of every restart, the cluster silhouettes, assignments and centroids."""
    fn = all_filename(resultdir, casefilename, X_shape[0], X_shape[1], nk, nNMF, ext)
    nan = np.float64(np.nan)
    items = {"W": details["W"], "H": details["H"], "Wmean": details["Wmean"], "Hmean": details["Hmean"],
             "Wvar": nan if details.get("Wvar") is None else details["Wvar"], "Hvar": nan if details.get("Hvar") is None else details["Hvar"],
             "Wbest": details["Wbest"], "Hbest": details["Hbest"], "fit": details["fit"],
             "Cluster Silhouettes": nan if details.get("clustersil") is None else np.asarray(details["clustersil"]).reshape(-1, 1),
             "Cluster assignments": nan if details.get("labels") is None else details["labels"],
             "Cluster centroids": nan if details.get("centroids") is None else details["centroids"]}
    _write(fn, **items)
    return fn


def load_all(X_shape, nk: int, nNMF: int, *, resultdir=".", casefilename="nmfk", ext=".npz"):
    """loadall=true (NMFkExecute.jl:492-506): -> (WBig (R,n,k), HBig (R,k,m), objvalue (R,)) or None when the file is missing."""
    fn = all_filename(resultdir, casefilename, X_shape[0], X_shape[1], nk, nNMF, ext)
    if not os.path.isfile(fn):
        return None
    return _read(fn, ("W", "H", "fit"))
