#!/usr/bin/env python3
"""bench.py - NMF restart-iterations/sec of the B200-native NMFk hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): synthetic nonnegative mixture X = W0*H0, 1000 x 200
Float64 (k0 = 5), NMFk.execute(X, 2:10, 100; method=:simple) with the reference's stop rule
(maxiter=10000, tolOF=1e-3, ...).  One step = one such execute over one X: 900 restarts
(9 values of k x 100), all solved concurrently on the device.

value  = restart-iterations / second of the solver phase with X and the initial factors already
         resident in HBM (device time, CUDA events on the library's launching stream).
e2e    = the same metric through the public API (nmfk_b200.execute == the C-ABI nmfk_execute)
         with HOST buffers: X and all initial factors are copied from pinned host memory and
         the per-k best factors / fit / robustness / aic / kopt are read back inside the timed
         region, which also contains clustering, silhouettes and selection.
N > 1  : restarts are sharded (weak scaling: every rank solves 100 restarts per k with its own
         seeds); H stacks and objectives are all-gathered over NCCL for the clustering of the
         R = 100*N solutions per k, which is split by k across ranks.

--impl reference times the CPU restatement of the reference (oracle/nmfk_oracle.py: NumPy +
OpenBLAS, the operation sequence of NMFkMultiplicative.jl) on the host cores - Julia itself is
not installed here or on the GPU box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "nmfk.jl_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "nmf_restart_iterations_per_sec"
UNIT = "restart-iterations/s"
N_ROWS, N_COLS, K0 = 1000, 200, 5
KS = list(range(2, 11))
R_PER_RANK = 100
SEED_X, SEED0 = 2015, 2015
WORKLOAD = "C2: synthetic mixture 1000x200 Float64 (k0=5), execute(X, 2:10, nNMF=100; method=:simple), maxiter=10000"


def algorithmic_flops(iters_by_k):
    """KL multiplicative update: 8*n*m*k flops per restart-iteration."""
    return float(sum(8.0 * N_ROWS * N_COLS * k * it for k, it in iters_by_k.items()))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(iters_per_k=100, ks=KS):
    """Bounded sample of the C2 workload on the host cores: `iters_per_k` iterations of ONE restart
    for every k of the sweep, literal reference operation sequence (oracle)."""
    from oracle import nmfk_oracle as o
    from nmfk_b200 import synth
    X = synth.mixture(N_ROWS, N_COLS, K0, seed=SEED_X)
    tot = 0
    t0 = time.perf_counter()
    for k in ks:
        W0, H0 = synth.philox_inits(SEED0, 1, N_ROWS, k, N_COLS)
        inf = {}
        o.nmf_multiplicative(X.copy(order="F"), k, Winit=W0[0], Hinit=H0[0], maxiter=iters_per_k, info=inf)
        tot += inf["iters"]
    dt = time.perf_counter() - t0
    return tot, dt


def bench_tiled(ctx, nb, synth, tag, n, m, k, R, dtype, k0, iters=20):
    """One of the large BASELINE configurations on ONE GPU through the tiled engine: a fixed number of iterations with
    every restart active; the tensor-core pass (tcgen05 kind::tf32 3-term split for Float32, DMMA m8n8k4 for Float64)
    next to the scalar-FMA pass kernel and a bounded CPU sample.  Reported under "also" (the headline stays C2)."""
    from oracle import nmfk_oracle as o
    f32 = dtype == np.float32
    X = synth.mixture(n, m, k0, seed=SEED_X, dtype=dtype)
    ctx.set_X(X)
    peak = max(ctx.measure_peak(2 if f32 else 1) for _ in range(2))
    out = {}
    for name, eng in (("tensor", 2), ("scalar_fma", 4)):
        best = None
        for rep, it in enumerate((3, iters, iters)):  # the first solve is the warm-up; best of two timed solves
            b = ctx.batch(k, R)
            b.init_random(SEED0)
            sampler = ClockSampler(0)
            sampler.start()
            ctx.solve([b], nb.default_params(maxiter=it, engine=eng))
            clk = sampler.stop()
            ms = ctx.last_solve_ms
            tot = int(b.get(factors=False)["iters"].sum())
            b.close()
            if rep > 0 and (best is None or ms < best[0]):
                best = (ms, tot, clk)
        ms, tot, clk = best
        tf = 8.0 * n * m * k * tot / ms / 1e9
        out[name] = {"value": tot / ms * 1e3, "unit": UNIT, "ms": ms, "restart_iterations": tot,
                     "algorithmic_tflops": tf, "frac_of_peak": tf / peak, "clocks": clk}
    W0, H0 = synth.philox_inits(SEED0, 1, n, k, m)
    inf = {}
    t0 = time.perf_counter()
    o.nmf_multiplicative(np.asfortranarray(X.astype(np.float64)), k, Winit=W0[0], Hinit=H0[0], maxiter=3, info=inf)
    dt = time.perf_counter() - t0
    return {"workload": "%s, %d iterations, all restarts active" % (tag, iters),
            "engine": "tiled; " + ("tcgen05.mma kind::tf32 3-term split (kl_tiled_tc.cu)" if f32 else "DMMA m8n8k4 (kl_tiled_dmma.cu)"),
            "peak": {"value": peak, "unit": "TFLOP/s", "what": "FP32 FFMA (own micro-benchmark)" if f32 else "FP64 DMMA (own micro-benchmark)"},
            "tensor_pass": out["tensor"], "scalar_fma_pass": out["scalar_fma"],
            "cpu_baseline": {"value": inf["iters"] / dt, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                             "sample": "3 iterations of 1 restart (oracle, Float64 like the reference)"}}


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_k = 60
    for _ in range(args.warmup):
        cpu_sample(10)
    tot, dt = 0, 0.0
    for _ in range(args.steps):
        t, d = cpu_sample(per_k)
        tot += t
        dt += d
    val = tot / dt
    cores = blas_threads()
    sample = "%d iterations of 1 restart for each k in 2:10 per step (%d restart-iterations/step)" % (per_k, per_k * len(KS))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of NMFk.jl (Julia unavailable): NumPy/OpenBLAS, "
                       "5 n*m*k products + 2 n*m divides per iteration, restarts serial"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary C3 measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import nmfk_b200 as nb
    from nmfk_b200 import dist as nbdist
    from nmfk_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if use_dist:
            td.barrier()
        torch.cuda.synchronize()

    ctx = nb.Context(local_rank)
    X = synth.mixture(N_ROWS, N_COLS, K0, seed=SEED_X)
    # pinned host copies of every input of one step (X + the initial factors of this rank's restarts)
    seed_rank = SEED0 + rank * R_PER_RANK
    Xpin = torch.from_numpy(X.T.copy()).pin_memory()  # (m, n) C-order == n x m column-major
    inits = {}
    h2d = Xpin.numel() * 8
    for k in KS:
        W0, H0 = synth.philox_inits(seed_rank, R_PER_RANK, N_ROWS, k, N_COLS)
        Wp = torch.from_numpy(np.ascontiguousarray(np.transpose(W0, (0, 2, 1)))).pin_memory()
        Hp = torch.from_numpy(np.ascontiguousarray(np.transpose(H0, (0, 2, 1)))).pin_memory()
        inits[k] = (Wp, Hp)
        h2d += (Wp.numel() + Hp.numel()) * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    peaks = {"fp64_dmma_tflops": max(ctx.measure_peak(1) for _ in range(2)),
             "fp64_dfma_tflops": max(ctx.measure_peak(0) for _ in range(2))}

    # ---------------- device-resident arm: `value` ----------------
    ctx.set_X(X)
    params = nb.default_params()

    def make_batches():
        bs = [ctx.batch(k, R_PER_RANK) for k in KS]
        for b in bs:
            b.init_random(seed_rank)  # same Philox streams as the pinned host inits, generated in HBM
        return bs

    def device_step():
        bs = make_batches()
        flush.fill_(1)
        torch.cuda.synchronize()
        launches0 = ctx.launches
        ctx.solve(bs, params)  # timed inside by CUDA events on the launching stream
        ms = ctx.last_solve_ms
        iters = {b.k: int(b.get(factors=False)["iters"].sum()) for b in bs}
        nl = ctx.launches - launches0
        for b in bs:
            b.close()
        return ms, iters, nl

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    tot_ms, tot_it, launches, iters_sum = 0.0, 0, 0, {k: 0 for k in KS}
    for _ in range(args.steps):
        ms, iters, nl = device_step()
        tot_ms += ms
        launches += nl
        for k, v in iters.items():
            iters_sum[k] += v
        tot_it += sum(iters.values())
    barrier()
    wall_dev = time.perf_counter() - t_wall0
    clocks = sampler.stop()

    # ---------------- end-to-end arm through the public API with host buffers ----------------
    def e2e_step():
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = nbdist.execute_sharded(ctx, Xpin.numpy().T, KS, R_PER_RANK,
                                     inits={k: (inits[k][0].numpy(), inits[k][1].numpy()) for k in KS},
                                     stack_layout=True, rank=rank, world=world)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, out

    for _ in range(min(args.warmup, 1)):
        e2e_step()
    barrier()
    e2e_t, e2e_it, d2h, kopt = 0.0, 0, 0, None
    for _ in range(args.steps):
        dt, out = e2e_step()
        e2e_t += dt
        e2e_it += out["total_iters_local"]
        d2h = out["d2h_bytes"]
        kopt = out["kopt"]
    barrier()

    # ---------------- reduce over ranks: max time, summed work ----------------
    if use_dist:
        t = torch.tensor([tot_ms, e2e_t], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        w = torch.tensor([float(tot_it), float(e2e_it), float(launches), algorithmic_flops(iters_sum)],
                         dtype=torch.float64, device="cuda")
        td.all_reduce(w, op=td.ReduceOp.SUM)
        tot_ms_max, e2e_t_max = t.tolist()
        tot_it_all, e2e_it_all, launches_all, flops_all = w.tolist()
    else:
        tot_ms_max, e2e_t_max = tot_ms, e2e_t
        tot_it_all, e2e_it_all, launches_all, flops_all = float(tot_it), float(e2e_it), float(launches), \
            algorithmic_flops(iters_sum)

    if rank == 0:
        value = tot_it_all / (tot_ms_max * 1e-3)
        achieved = flops_all / world / (tot_ms_max * 1e-3) / 1e12  # per GPU
        peak = peaks["fp64_dmma_tflops"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "restarts_per_gpu_per_k": R_PER_RANK, "engine": "resident (one CTA per restart)",
                       "l2": "256 MiB written between steps (X itself is 1.6 MB and L2-resident by design)",
                       "kopt": kopt, "restart_iterations_per_step": tot_it_all / args.steps,
                       "wall_s_device_arm": wall_dev},
            "clocks": clocks,
            "e2e": {"value": e2e_it_all / e2e_t_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_t_max / args.steps * 1e3},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": 17534720,
                         "traffic_note": "dram__bytes_read + dram__bytes_write of one ncu --set full capture of "
                                         "kl_resident_dmma_kernel<10> (30 iterations of 148 restarts, "
                                         "profiles/r01_resident_dmma_v3_k10.txt): X (1.6 MB, both orientations) and the "
                                         "factors are read from DRAM once and stay in L2 / shared memory - the kernel is "
                                         "FP64-pipe and L2-latency bound, not DRAM bound",
                         "kernel": "kl_resident_dmma_kernel<K,false> (9 instantiations, one per k, concurrent streams; "
                                   "DMMA m8n8k4 + DFMA remainder columns share the FP64 pipe)",
                         "note": "FP64 work: MEASURED_PEAKS.json holds only HBM and bf16 peaks, so the denominator is the "
                                 "FP64 DMMA (mma.sync m8n8k4) throughput measured by nmfk_measure_peak in this run; the "
                                 "DFMA pipe measured %.1f TFLOP/s. achieved = 8*n*m*k flops per restart-iteration / "
                                 "event time of the solve" % peaks["fp64_dfma_tflops"]},
        }
        if not args.no_cpu_baseline and world == 1:
            tot, dt = cpu_sample(60)
            line["cpu_baseline"] = {"value": tot / dt, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                                    "sample": "60 iterations of 1 restart for each k in 2:10 (540 restart-iterations), "
                                              "oracle/nmfk_oracle.py (NumPy/OpenBLAS restatement of NMFkMultiplicative.jl)"}
        if not args.no_also and world == 1:
            line["also"] = {}
            for key, cfg in (("C3", ("C3: synthetic mixture 10000x10000 Float32, k=16, nNMF=64", 10000, 10000, 16, 64, np.float32, 16)),
                             ("C4_k32", ("C4 at k=32 on one GPU: synthetic mixture 100000x2000 Float64, 32 of the 256 restarts",
                                         100000, 2000, 32, 32, np.float64, 8))):
                try:
                    line["also"][key] = bench_tiled(ctx, nb, synth, *cfg)
                except Exception as e:  # the headline must survive a failure of a secondary measurement
                    line["also"][key] = {"error": repr(e)}
        print(json.dumps(line))
    ctx.close()
    if use_dist:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
