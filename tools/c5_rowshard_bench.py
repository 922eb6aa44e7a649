"""BASELINE config C5 (2,000,000 x 1000 Float32, k = 24, nNMF = 32) row-sharded over the GPUs of one box:
every rank holds n_global / world rows of X and W, H is replicated, the tiled engine (tcgen05 pass kernel) all-reduces
colsum(W) and the k x m x R numerators once per iteration over NCCL.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/c5_rowshard_bench.py [n_global] [iters]
-> one JSON line from rank 0 (restart-iterations/s = R * iters / max-over-ranks device time of the solve)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nmfk.jl_b200", "python")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import dist as nbdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n_global = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m, k, R, k0 = 1000, 24, 32, 24
torch.cuda.set_device(local)
td.init_process_group("gloo")
r0, r1 = nbdist.row_block(n_global, rank, world)
# this rank's rows of the synthetic mixture: H0 from the common stream, W0 rows from a per-rank stream
H0 = np.random.Generator(np.random.Philox(key=2015)).random((k0, m))
W0 = np.random.Generator(np.random.Philox(key=2016 + rank)).random((r1 - r0, k0))
Xloc = np.asfortranarray((W0 @ H0).astype(np.float32))
del W0
out = None
for it in (3, iters, iters):  # warm-up, then best of two
    with nb.Context(local) as ctx:
        o = nbdist.solve_rowsharded(ctx, Xloc, k, R, rank=rank, world=world, n_global=n_global,
                                    unique_id=nbdist.exchange_unique_id(rank), seed0=2015, params=nb.default_params(maxiter=it))
    o = {key: o[key] for key in ("iters", "obj_norm", "solve_ms")}
    if it == iters and (out is None or o["solve_ms"] < out["solve_ms"]):
        out = o
ms = torch.tensor([out["solve_ms"]], dtype=torch.float64)
td.all_reduce(ms, op=td.ReduceOp.MAX)
if rank == 0:
    tot = int(out["iters"].sum())
    flops = 8.0 * n_global * m * k * tot
    print(json.dumps(dict(workload="C5: %dx%d Float32, k=%d, nNMF=%d, X row-sharded over %d GPU(s), %d iterations" % (n_global, m, k, R, world, iters),
                          n_gpus=world, ms=float(ms[0]), restart_iterations=tot, value=tot / float(ms[0]) * 1e3, unit="restart-iterations/s",
                          algorithmic_tflops_total=flops / float(ms[0]) / 1e9, obj_norm_first=float(out["obj_norm"][0]))))
td.barrier()
td.destroy_process_group()
