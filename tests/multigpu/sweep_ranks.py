"""Restart-sharded execute on N GPUs (one process per GPU, torchrun) through nmfk_sweep: the library's own NCCL communicator
gathers the H stacks / restart states and broadcasts the best W; the result on EVERY rank must equal the CPU oracle's
execute(X, ks, nNMF = N * R_local) started from the same Philox streams - labels-level decisions included (robustness is the
minimum cluster silhouette of the gathered solutions).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu/sweep_ranks.py"""
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
from oracle import nmfk_oracle as o  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("gloo")  # carries the 128-byte NCCL id only
rel = lambda a, b_: float(np.max(np.abs(np.asarray(a) - np.asarray(b_))) / max(np.max(np.abs(b_)), 1e-300))

# 1) the README problem: Float64, resident engine, full stop rule
X = synth.readme_bss()
ks, R_local, seed = [2, 3, 4], 3, 100
with nb.Context(local) as ctx:
    out = nbdist.execute_sharded(ctx, X, ks, R_local, seed0=seed, rank=rank, world=world)
Wo, Ho, fito, robo, aico, kopto = o.execute(X.copy(), ks, world * R_local, seed=seed)
assert out["kopt"] == kopto, (out["kopt"], kopto)
for i, k in enumerate(ks):
    assert abs(out["robustness"][i] - robo[k - 1]) < 1e-7, (k, out["robustness"][i], robo[k - 1])
    assert abs(out["fit"][i] - fito[k - 1]) <= 1e-6 * max(fito[k - 1], 1e-6), (k, out["fit"][i], fito[k - 1])
    assert abs(out["aic"][i] - aico[k - 1]) <= 1e-5 * abs(aico[k - 1])
    assert rel(out["W"][k], Wo[k]) < 1e-6 and rel(out["H"][k], Ho[k]) < 1e-6, k
assert out["total_iters"] > out["total_iters_local"] > 0

# 2) explicit initial factors per rank, tiled engine (DMMA), fixed iteration budget: cluster labels of the gathered
#    world * R_local solutions against the oracle's clustersolutions on the same solutions
n, m, k, R2 = 600, 120, 4, 4
X2 = synth.mixture(n, m, 4, seed=3)
W0, H0 = synth.philox_inits(500, world * R2, n, k, m)
sl = slice(rank * R2, (rank + 1) * R2)
with nb.Context(local) as ctx:
    out2 = nbdist.execute_sharded(ctx, X2, [k], R2, inits={k: (W0[sl], H0[sl])}, rank=rank, world=world,
                                  params=nb.default_params(maxiter=150, engine=2))
W2, H2, f2, r2, a2 = o.execute_k(X2.copy(), k, world * R2, inits=[(W0[i].copy(), H0[i].copy()) for i in range(world * R2)], maxiter=150)
assert abs(out2["robustness"][0] - r2) < 1e-7 and abs(out2["fit"][0] - f2) <= 1e-7 * f2
assert rel(out2["W"][k], W2) < 1e-7 and rel(out2["H"][k], H2) < 1e-7

# every rank holds the same answer
box = [None] * world
td.all_gather_object(box, (out["kopt"], out["robustness"].tobytes(), out2["W"][k].tobytes()))
assert all(b == box[0] for b in box)
td.barrier()
if rank == 0:
    print("SWEEP OK world=%d kopt=%s robustness=%s total_iters=%d (local %d)" % (world, out["kopt"], np.round(out["robustness"], 6).tolist(),
                                                                               out["total_iters"], out["total_iters_local"]))
td.destroy_process_group()
