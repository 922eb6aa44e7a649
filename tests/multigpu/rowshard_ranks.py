"""Row-sharded NMFmultiplicative on N GPUs (one process per GPU, torchrun): every rank holds a block of
rows of X; the result must match the single-GPU solve of the whole matrix (same engine) and the oracle.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu/rowshard_ranks.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "nmfk.jl_b200", "python")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import dist as nbdist  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("gloo")  # side channel for the 128-byte id only; the data path is the library's own NCCL communicator
n, m, k, R = 1501, 96, 6, 5
X = synth.mixture(n, m, 4, seed=9)
W0, H0 = synth.philox_inits(33, R, n, k, m)
r0, r1 = nbdist.row_block(n, rank, world)
uid = nbdist.exchange_unique_id(rank)
params = nb.default_params(maxiter=400)
with nb.Context(local) as ctx:
    out = nbdist.solve_rowsharded(ctx, X[r0:r1], k, R, rank=rank, world=world, n_global=n, unique_id=uid,
                                  Winit_local=W0[:, r0:r1, :], Hinit=H0, params=params)
# single-GPU solve of the whole matrix with the same (tiled) engine
with nb.Context(local) as ctx1:
    ctx1.set_X(X)
    b = ctx1.batch(k, R)
    b.set_init(W0, H0)
    ctx1.solve([b], nb.default_params(maxiter=400, engine=2))
    ref = b.get()
    b.close()
rel = lambda a, b_: float(np.max(np.abs(a - b_)) / np.max(np.abs(b_)))
assert np.array_equal(out["iters"], ref["iters"]), (out["iters"], ref["iters"])
assert np.array_equal(out["stop_reason"], ref["stop_reason"])
eH, eW = rel(out["H"], ref["H"]), rel(out["W_local"], ref["W"][:, r0:r1, :])
eo = rel(out["obj_norm"], ref["obj_norm"])
assert eH < 1e-9 and eW < 1e-9 and eo < 1e-9, (eH, eW, eo)
# H replicated bit for bit
Hs = [None] * world
td.all_gather_object(Hs, out["H"].tobytes())
assert all(h == Hs[0] for h in Hs)
# device-generated initial factors: rows of the global Philox streams
with nb.Context(local) as ctx2:
    o2 = nbdist.solve_rowsharded(ctx2, X[r0:r1], k, R, rank=rank, world=world, n_global=n, unique_id=nbdist.exchange_unique_id(rank),
                                 seed0=33, params=params)
assert np.array_equal(o2["iters"], out["iters"]) and rel(o2["H"], out["H"]) < 1e-12
# Float32: every rank runs the tcgen05 pass kernel on its rows (n_local % 4 == 0); numerators all-reduced in Float32
n32, m32, k32, R32 = 2048 * world, 256, 16, 5
X32 = synth.mixture(n32, m32, 4, seed=11, dtype=np.float32)
W32, H32 = synth.philox_inits(44, R32, n32, k32, m32, dtype=np.float32)
q0, q1 = nbdist.row_block(n32, rank, world)
p32 = nb.default_params(maxiter=30)
with nb.Context(local) as ctx3:
    o32 = nbdist.solve_rowsharded(ctx3, X32[q0:q1], k32, R32, rank=rank, world=world, n_global=n32,
                                  unique_id=nbdist.exchange_unique_id(rank), Winit_local=W32[:, q0:q1, :], Hinit=H32, params=p32)
with nb.Context(local) as ctx4:
    ctx4.set_X(X32)
    b = ctx4.batch(k32, R32)
    b.set_init(W32, H32)
    ctx4.solve([b], nb.default_params(maxiter=30, engine=2))
    r32 = b.get()
    b.close()
e32H, e32W = rel(o32["H"], r32["H"].astype(np.float64)), rel(o32["W_local"], r32["W"][:, q0:q1, :].astype(np.float64))
assert np.array_equal(o32["iters"], r32["iters"]) and e32H < 1e-4 and e32W < 1e-4, (e32H, e32W)
td.barrier()
if rank == 0:
    print("ROWSHARD F32 (tcgen05) OK relerr H=%.2e W=%.2e solve_ms=%.2f" % (e32H, e32W, o32["solve_ms"]))
    print("ROWSHARD OK world=%d iters=%s relerr H=%.2e W=%.2e obj=%.2e solve_ms=%.2f" % (world, out["iters"].tolist(), eH, eW, eo, out["solve_ms"]))
td.destroy_process_group()
