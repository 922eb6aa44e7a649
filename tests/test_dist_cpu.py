"""Host-side logic of the restart-sharded (multi-GPU) path, world_size 2 over gloo on CPU:
rank-major gathering of H stacks / objectives / iteration counts, the global<->(rank, restart)
index maps, ownership of k values and the final merge + kopt selection.  (The device kernels
are covered by the -m gpu tests; nothing here computes a factorization.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from nmfk_b200 import dist as nbdist
from oracle import nmfk_oracle as o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R_local, m, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        H = torch.from_numpy(rng.random((R_local, m, k)))
        obj = torch.from_numpy(rng.random(R_local))
        its = torch.from_numpy(rng.integers(10, 1000, R_local).astype(np.int32))
        Hall, oall, iall = nbdist.gather_solutions(H, obj, its)
        q.put((rank, Hall.numpy(), oall.numpy(), iall.numpy()))
    finally:
        td.destroy_process_group()


def test_gather_solutions_rank_major_gloo():
    world, R_local, m, k = 2, 3, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R_local, m, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Hexp, oexp, iexp = [], [], []
    for r in range(world):
        rng = np.random.default_rng(100 + r)
        Hexp.append(rng.random((R_local, m, k)))
        oexp.append(rng.random(R_local))
        iexp.append(rng.integers(10, 1000, R_local).astype(np.int32))
    Hexp, oexp, iexp = np.concatenate(Hexp), np.concatenate(oexp), np.concatenate(iexp)
    for rank, Hall, oall, iall in got:
        assert np.array_equal(Hall, Hexp) and np.array_equal(oall, oexp) and np.array_equal(iall, iexp)
    # the sort the owner then performs is sortperm over the rank-major vector
    order = np.argsort(oexp, kind="stable")
    assert nbdist.split_global(int(order[0]), R_local) == divmod(int(order[0]), R_local)


def test_index_maps_and_ownership():
    for world in (1, 2, 4, 8):
        R = 7
        seen = set()
        for rank in range(world):
            for r in range(R):
                g = nbdist.global_index(rank, r, R)
                assert nbdist.split_global(g, R) == (rank, r)
                seen.add(g)
        assert seen == set(range(world * R))
        owners = [nbdist.owner_of(i, world) for i in range(9)]
        assert set(owners) == set(range(min(world, 9))) and max(owners) < world
        # balanced: no rank owns more than ceil(9/world) values of k
        assert max(owners.count(r) for r in range(world)) == -(-9 // world)


def test_merge_sweep_and_aic_match_oracle_rules():
    ks = [2, 3, 4, 5]
    per_k = {2: dict(fit=3.7, robustness=0.98, aic=-145.0), 3: dict(fit=0.01, robustness=0.85, aic=-550.0),
             4: dict(fit=0.001, robustness=0.26, aic=-713.0), 5: dict(fit=0.003, robustness=-0.5, aic=-570.0)}
    fit, rob, aic, kopt = nbdist.merge_sweep(ks, per_k)
    assert kopt == o.getk(ks, [per_k[k]["robustness"] for k in ks]) == 3
    assert fit[0] == np.inf and rob[0] == -1 and fit[2] == 0.01 and aic[4] == -570.0
    for k in ks:
        per_k[k]["fit"] = np.inf
    assert nbdist.merge_sweep(ks, per_k)[3] == 0  # "No successful NMFk runs" (NMFkExecute.jl:206-208)
    # aic formula of NMFkExecute.jl:697-708
    assert nbdist.aic(15, 5, 2, 0, 3.733236078245728) == pytest.approx(-145.01595060344386, rel=1e-12)
    assert nbdist.aic(15, 5, 2, 5, 2.0) == pytest.approx(2 * (30 + 10) + 70 * np.log(2.0 / 70))


# ---------------------------------------------------------------------------------------------
# row-sharded X (BASELINE C5; NMFmultiplicative(::DArray), NMFkMultiplicative.jl:129-197): the exchange
# the tiled engine performs - sum-all-reduce of the k x m numerators and of colsum(W), W-update local -
# restated in NumPy over gloo must reproduce the dense iteration of the oracle.
# ---------------------------------------------------------------------------------------------
def test_row_block_partition():
    for n in (1, 7, 10, 1000, 2_000_000):
        for world in (1, 2, 3, 8):
            blocks = [nbdist.row_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _rowshard_worker(rank, world, port, n, m, k, niter, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        X = rng.random((n, 3)) @ rng.random((3, m))
        W = rng.random((n, k))
        H = rng.random((k, m))
        r0, r1 = nbdist.row_block(n, rank, world)
        Xl, Wl = X[r0:r1], W[r0:r1].copy()
        for _ in range(niter):
            num = Wl.T @ (Xl / (Wl @ H))          # k x m numerators of this rank's rows
            den = Wl.sum(axis=0)                  # colsum(W) of this rank's rows
            buf = torch.from_numpy(np.concatenate([den, num.ravel()]))
            td.all_reduce(buf)                    # the one exchange of an iteration
            den, num = buf.numpy()[:k], buf.numpy()[k:].reshape(k, m)
            H = H * num / den[:, None]            # :67, identical on every rank
            Wl = Wl * ((Xl / (Wl @ H)) @ H.T) / H.sum(axis=1)[None, :]   # :70, rows are independent
        ssq = torch.tensor([float(np.sum((Xl - Wl @ H) ** 2))], dtype=torch.float64)
        td.all_reduce(ssq)                        # :74 objective = sum over all rows
        q.put((rank, r0, r1, Wl, H, float(ssq[0])))
    finally:
        td.destroy_process_group()


def test_rowsharded_exchange_reproduces_dense_iteration_gloo():
    world, n, m, k, niter = 2, 37, 9, 3, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rowshard_worker, args=(r, world, port, n, m, k, niter, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    X = rng.random((n, 3)) @ rng.random((3, m))
    W0 = rng.random((n, k))
    H0 = rng.random((k, m))
    Ws, Hs = [], []
    o.nmf_multiplicative(np.asfortranarray(X.copy()), k, Winit=W0, Hinit=H0, maxiter=niter,
                         trace=lambda it, W, H, obj: (Ws.append(W.copy()), Hs.append(H.copy())))
    Wd, Hd = Ws[niter - 1], Hs[niter - 1]
    Wsh = np.concatenate([g[3] for g in got])
    assert [g[1:3] for g in got] == [nbdist.row_block(n, r, world) for r in range(world)]
    assert np.allclose(got[0][4], got[1][4], rtol=0, atol=0)          # H replicated bit for bit
    assert np.max(np.abs(got[0][4] - Hd)) <= 1e-12 * np.max(Hd)
    assert np.max(np.abs(Wsh - Wd)) <= 1e-12 * np.max(Wd)
    assert got[0][5] == got[1][5] == pytest.approx(float(np.sum((X - Wd @ Hd) ** 2)), rel=1e-10)
