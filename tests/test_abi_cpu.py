"""CPU-only checks of the C-ABI boundary: the library builds/loads, exports every symbol that
include/nmfk_b200.h declares, fails loudly without a GPU, and its pure-host entry points
(getk, signalorder, Philox stream) agree with the oracle / NumPy."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nmfk_b200
from nmfk_b200 import _lib
from oracle import nmfk_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nmfk_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nmfk_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert declared == set(_lib.SIGNATURES), "python binding and header disagree"
    assert lib.nmfk_abi_version() == 2


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nmfk_b200.NMFkError) as ei:
        nmfk_b200.Context()
    assert ei.value.status == -8 and "no CPU fallback" in str(ei.value)


def test_default_params_are_the_reference_defaults():
    p = nmfk_b200.default_params()
    # NMFkExecute.jl:729 (maxiter, tol) and NMFkMultiplicative.jl:24 (the rest)
    assert (p.tol, p.tolOF, p.maxiter, p.maxbaditers, p.maxreattempts, p.stopconv, p.check_every) == \
        (1e-19, 1e-3, 10000, 10, 2, 1000, 10)
    assert p.eps_clamp == np.finfo(np.float64).eps and p.weight == 1.0 and p.normalize == 1
    assert C.sizeof(_lib.Params) == 4 * 8 + 14 * 4 and C.sizeof(_lib.XInfo) == 56


def test_philox_stream_is_numpy_philox():
    lib = _lib.load()
    for seed in (0, 1, 2016, 2 ** 40 + 7):
        out = np.empty(1001)
        assert lib.nmfk_philox_host(seed, 1001, out.ctypes.data_as(_lib._pdbl)) == 0
        ref = np.random.Generator(np.random.Philox(key=seed)).random(1001)
        assert np.array_equal(out, ref)


def test_getk_matches_oracle():
    rng = np.random.default_rng(0)
    cases = [([2, 3, 4, 5], [0.99, 0.85, -0.57, -0.67]), ([2, 3, 4], [0.1, 0.2, 0.3]), ([2, 3], [np.nan, np.nan]),
             ([3], [0.6]), ([3], [0.4]), ([2, 3, 4], [0.9, 0.2, 0.7]), ([2, 3, 4], [np.nan, 0.7, 0.1])]
    for _ in range(50):
        ks = list(range(2, 2 + int(rng.integers(1, 8))))
        rb = rng.uniform(-1, 1, len(ks))
        rb[rng.random(len(ks)) < 0.15] = np.nan
        cases.append((ks, list(rb)))
    for ks, rb in cases:
        for strict in (True, False):
            assert nmfk_b200.getk(ks, rb, 0.5, strict) == o.getk(ks, rb, 0.5, strict), (ks, rb, strict)


def test_signalorder_matches_oracle():
    rng = np.random.default_rng(1)
    for dt in (np.float64, np.float32):
        for _ in range(10):
            n, k, m = int(rng.integers(3, 40)), int(rng.integers(1, 9)), int(rng.integers(2, 30))
            W = rng.random((n, k)).astype(dt)
            H = rng.random((k, m)).astype(dt)
            assert list(nmfk_b200.signalorder(W, H)) == list(o.signalorder(W.astype(np.float64), H.astype(np.float64)))
    W = np.array([[1.0, 1.0, 2.0], [1.0, 1.0, 2.0]])
    H = np.array([[1.0, 1.0], [1.0, 1.0], [3.0, 3.0]])
    assert list(nmfk_b200.signalorder(W, H)) == [2, 0, 1]  # ties keep their order (stable sortperm)


def test_julia_shim_ccalls_match_the_header():
    """The Julia shim cannot run here (no Julia): at least every `ccall` names an exported entry point and passes as many
    arguments as the header declares (a renamed or re-shaped entry point would otherwise only fail on a user's machine)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "nmfk_b200.h")).read()
    protos = {}
    for mm in re.finditer(r"\b(?:int32_t|int64_t|const char\*|void|double)\s+(nmfk_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        args = mm.group(2).strip()
        protos[mm.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    jl = open(os.path.join(root, "nmfk.jl_b200", "julia", "NMFkB200", "src", "NMFkB200.jl")).read()
    calls = list(re.finditer(r"ccall\(\(:(nmfk_\w+),\s*\w+\),\s*\w+,\s*\(([^)]*)\)", jl, re.S))
    assert len(calls) >= 20
    for mm in calls:
        name, types = mm.group(1), mm.group(2).strip()
        n = 0 if types == "" else len([t for t in types.split(",") if t.strip()])
        assert name in protos, name
        assert protos[name] == n, (name, protos[name], n)
