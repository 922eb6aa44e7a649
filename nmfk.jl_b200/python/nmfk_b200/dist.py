"""Restart-sharded execute across the GPUs of one box: one process per GPU, torch.distributed
(NCCL) for the plumbing.

The reference farms restarts out with `Distributed.pmap` (NMFkExecute.jl:511-526) and gathers the
(W, H, objvalue) tuples by value.  Here every rank solves its own restarts for every k (X is
replicated: it is small next to the factor stacks), then only what the robustness analysis needs
crosses NVLink: the H stacks (k x m per restart) and the restart states are all-gathered, the
clustering + silhouettes of the R_total = world * R_local solutions of a given k run on the rank
that owns that k (k index mod world), and the best restart's W (n x k) is broadcast from the rank
that solved it.  There is no collective inside the iteration loop.  All of this lives behind the
C ABI (nmfk_sweep, library-owned NCCL communicator); torch.distributed only carries the 128-byte
NCCL id.

The helpers that do not touch the device (`row_block`, `owner_of`, `global_index`, `split_global`,
`gather_solutions`, `merge_sweep`) are covered by world_size-2 gloo tests on CPU
(tests/test_dist_cpu.py)."""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np

from . import api
from ._lib import check


def row_block(n: int, rank: int, world: int):
    """Rows [r0, r1) of an n-row matrix owned by `rank` (contiguous blocks, sizes differ by at most 1)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def exchange_unique_id(rank: int, group=None) -> bytes:
    """NCCL unique id of the library's own communicator: made on rank 0, broadcast with torch.distributed
    (any backend: 128 bytes on the host)."""
    import torch.distributed as td

    box = [api.comm_unique_id() if rank == 0 else None]
    td.broadcast_object_list(box, src=0, group=group)
    return box[0]


def solve_rowsharded(ctx, X_local, k: int, R: int, *, rank: int, world: int, n_global: int, unique_id=None,
                     Winit_local=None, Hinit=None, seed0: Optional[int] = None, params=None):
    """NMFmultiplicative for R restarts at one k on a matrix whose ROWS are split over `world` GPUs
    (one process per GPU; reference analogue: NMFmultiplicative(::DArray), NMFkMultiplicative.jl:129-197,
    with the full stop rule of the dense method).

    X_local: this rank's rows (row_block(n_global, rank, world)); Winit_local (R, n_local, k) and Hinit
    (R, k, m, identical on all ranks) or seed0 for the device Philox streams of the global matrices.
    Returns dict(W_local (R, n_local, k), H (R, k, m), obj_norm, obj_ssq, iters, stop_reason): everything
    but W_local is identical on all ranks."""
    r0, r1 = row_block(n_global, rank, world)
    assert X_local.shape[0] == r1 - r0, "X_local must hold rows row_block(n_global, rank, world)"
    ctx.comm_init(world, rank, unique_id, r0, n_global)
    ctx.set_X(X_local)
    p = params or api.default_params()
    p.engine = 2  # NMFK_ENGINE_TILED: the engine that exchanges partial sums
    b = ctx.batch(k, R)
    try:
        if Winit_local is not None:
            b.set_init(Winit_local, Hinit)
        else:
            b.init_random(int(seed0 or 0))
        ctx.solve([b], p)
        out = b.get()
    finally:
        b.close()
    return dict(W_local=out["W"], H=out["H"], obj_norm=out["obj_norm"], obj_ssq=out["obj_ssq"], iters=out["iters"],
                stop_reason=out["stop_reason"], solve_ms=ctx.last_solve_ms)


def owner_of(k_index: int, world: int) -> int:
    """Rank that clusters the solutions of the k_index-th entry of the sweep."""
    return k_index % world


def global_index(rank: int, r_local: int, R_local: int) -> int:
    """Position of restart r_local of `rank` in the gathered (rank-major) stack."""
    return rank * R_local + r_local


def split_global(g: int, R_local: int):
    return divmod(g, R_local)  # (rank, r_local)


def gather_solutions(H_local, obj_local, iters_local, group=None):
    """all_gather of one k's solutions.  H_local: tensor (R_local, m, k) (stack layout, any device),
    obj_local: (R_local,) float64, iters_local: (R_local,) int32.  Returns rank-major concatenations."""
    import torch
    import torch.distributed as td

    world = td.get_world_size(group)
    Hs = [torch.empty_like(H_local) for _ in range(world)]
    td.all_gather(Hs, H_local.contiguous(), group=group)
    meta_local = torch.stack([obj_local.to(torch.float64), iters_local.to(torch.float64)], dim=0).contiguous()
    metas = [torch.empty_like(meta_local) for _ in range(world)]
    td.all_gather(metas, meta_local, group=group)
    H = torch.cat(Hs, dim=0)
    obj = torch.cat([mt[0] for mt in metas])
    iters = torch.cat([mt[1] for mt in metas]).to(torch.int32)
    return H, obj, iters


def aic(n: int, m: int, k: int, nnan: int, phi: float) -> float:
    """aic = 2*numparameters + numobservations*log(phi/numobservations) (NMFkExecute.jl:697-708)."""
    nobs = n * m - nnan
    return 2.0 * (n * k + k * m) + nobs * math.log(phi / nobs) if phi > 0 else -math.inf


def merge_sweep(ks: Sequence[int], per_k: Dict[int, dict], cutoff: float = 0.5):
    """Assemble execute's return values (NMFkExecute.jl:196-232) from per-k results."""
    maxk = max(ks)
    fit = np.zeros(maxk)
    rob = np.zeros(maxk)
    aicv = np.zeros(maxk)
    fit[0] = np.inf
    rob[0] = -1
    for k in ks:
        fit[k - 1], rob[k - 1], aicv[k - 1] = per_k[k]["fit"], per_k[k]["robustness"], per_k[k]["aic"]
    idx = np.asarray(list(ks)) - 1
    kopt = 0 if np.all(np.isinf(fit[idx])) else api.getk(list(ks), rob[idx], cutoff)
    return fit, rob, aicv, kopt


def sweep_comm_init(ctx, rank: int, world: int, group=None):
    """The library's own NCCL communicator for the restart-sharded sweep (nmfk_ctx_sweep_comm_init): the 128-byte id is made on
    rank 0 and broadcast with torch.distributed (any backend) - the only thing the side channel carries."""
    if getattr(ctx, "_sweep_world", None) == world:
        return
    uid = exchange_unique_id(rank, group) if world > 1 else None
    import ctypes as C
    buf = C.create_string_buffer(bytes(uid), 128) if uid is not None else None
    check(ctx._lib.nmfk_ctx_sweep_comm_init(ctx._h, int(world), int(rank), buf), ctx._h)
    ctx._sweep_world = world


def execute_sharded(ctx, X, ks: Sequence[int], R_local: int, *, inits=None, seed0: Optional[int] = None,
                    stack_layout: bool = False, rank: int = 0, world: int = 1, cutoff: float = 0.5, params=None):
    """execute(X, ks, nNMF = world * R_local) with the restarts sharded over `world` ranks (one process per GPU): a thin
    binding of nmfk_sweep - solve, NCCL exchange (the library's own communicator), owner-side clustering and the selection
    of kopt all happen behind the C ABI, exactly as a Julia host would reach them with `ccall`.

    inits[k] = (Winit, Hinit) for THIS rank's restarts ((R,n,k)/(R,k,m), or with stack_layout=True already (R,k,n)/(R,m,k)
    C-contiguous == column-major stacks; either may be None); otherwise device Philox streams keyed seed0 + global restart
    number.  Every rank returns the same dict (W, H by k, fit, robustness, aic, kopt, total_iters, total_iters_local)."""
    import ctypes as C
    from . import _lib

    ks = [int(k) for k in ks]
    ctx.set_X(X)
    sweep_comm_init(ctx, rank, world)
    p = params or api.default_params()
    n, m, dt = ctx.n, ctx.m, ctx.np_dtype
    nks = len(ks)
    Wo = [np.empty((k, n), dtype=dt) for k in ks]
    Ho = [np.empty((m, k), dtype=dt) for k in ks]
    Wop = (C.c_void_p * nks)(*[w.ctypes.data for w in Wo])
    Hop = (C.c_void_p * nks)(*[h.ctypes.data for h in Ho])
    Wip = Hip = None
    keep = []
    if inits is not None:
        for k in ks:
            Wi, Hi = inits[k]
            if not stack_layout:
                Wi = None if Wi is None else np.ascontiguousarray(np.transpose(np.asarray(Wi, dtype=dt), (0, 2, 1)))
                Hi = None if Hi is None else np.ascontiguousarray(np.transpose(np.asarray(Hi, dtype=dt), (0, 2, 1)))
            keep.append((Wi, Hi))
        Wip = (C.c_void_p * nks)(*[None if w is None else w.ctypes.data for w, _ in keep])
        Hip = (C.c_void_p * nks)(*[None if h is None else h.ctypes.data for _, h in keep])
    fit, rob, aicv = np.empty(nks), np.empty(nks), np.empty(nks)
    kopt, tot, tot_local = C.c_int32(), C.c_int64(), C.c_int64()
    karr = np.asarray(ks, dtype=np.int32)
    check(ctx._lib.nmfk_sweep(ctx._h, karr.ctypes.data_as(_lib._pi32), nks, R_local, Wip, Hip, int(seed0 or 0), C.byref(p),
                              cutoff, Wop, Hop, fit.ctypes.data_as(_lib._pdbl), rob.ctypes.data_as(_lib._pdbl),
                              aicv.ctypes.data_as(_lib._pdbl), C.byref(kopt), C.byref(tot), C.byref(tot_local)), ctx._h)
    d2h = sum(w.nbytes + h.nbytes for w, h in zip(Wo, Ho)) + 3 * 8 * nks + 12
    return dict(W={k: w.T for k, w in zip(ks, Wo)}, H={k: h.T for k, h in zip(ks, Ho)}, fit=fit, robustness=rob,
                aic=aicv, kopt=(None if kopt.value < 0 else kopt.value), total_iters=int(tot.value),
                total_iters_local=int(tot_local.value), d2h_bytes=d2h)
