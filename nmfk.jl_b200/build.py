#!/usr/bin/env python3
"""Builds libnmfk_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python nmfk.jl_b200/build.py [--force]

Every translation unit is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
(no other architectures, no PTX fallback, no CPU path) and linked into
nmfk.jl_b200/lib/libnmfk_b200.so.  The .so is git-ignored but travels to the GPU box.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libnmfk_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
UNITS = ["capi.cu", "kl_resident_f64.cu", "kl_resident_f32.cu", "kl_dmma.cu", "kl_tiled.cu", "kl_tiled_f64.cu", "kl_tiled_f32.cu", "preprocess.cu", "cluster.cu",
         "objective.cu", "microbench.cu", "tc_selftest.cu", "kl_tiled_tc.cu", "kl_tiled_tc2.cu", "kl_tiled_dmma.cu", "fro_gemm.cu", "fro_gemm_f64.cu",
         "fro_solve.cu", "kmeans.cu", "sparsity.cu"]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "nmfk_b200.h"))
    return hs


def _compile(unit, force):
    src = os.path.join(SRC, unit)
    obj = os.path.join(OBJ, unit + ".o")
    stamp = obj + ".sha"
    dig = _digest([src] + _headers())
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return unit, False, ""
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return unit, True, r.stderr


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    rebuilt = False
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for unit, did, err in ex.map(lambda u: _compile(u, force), UNITS):
            rebuilt |= did
            if verbose and did:
                print("[nmfk build] compiled", unit, file=sys.stderr)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + [os.path.join(OBJ, u + ".o") for u in UNITS] + \
              ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[nmfk build] linked", LIB, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
