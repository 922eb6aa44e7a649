"""Does the clock sampler stall the solve?  12 C3 steps (20 iterations, 64 restarts) per condition: no sampler, `nvidia-smi -lms 200`
(the recipe's line), NVML in-process thread.  usage: stall_probe.py"""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import numpy as np  # noqa: E402
import nmfk_b200 as nb  # noqa: E402
from nmfk_b200 import synth  # noqa: E402

X = synth.mixture(10000, 10000, 16, seed=2015, dtype=np.float32)
Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


def steps(ctx, n):
    out = []
    for i in range(n):
        b = ctx.batch(16, 64)
        b.init_random(2015 + i)
        ctx.solve([b], nb.default_params(maxiter=20, engine=2))
        out.append(round(ctx.last_solve_ms, 1))
        b.close()
    return out


with nb.Context(0) as ctx:
    ctx.set_X(X)
    steps(ctx, 3)
    print("no sampler      ", steps(ctx, 12), flush=True)
    proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", "200"],
                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    time.sleep(1.0)
    print("nvidia-smi -lms ", steps(ctx, 12), flush=True)
    proc.terminate()
    time.sleep(0.5)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        stop = False
        rows = []

        def loop():
            while not stop:
                rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                             if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                time.sleep(0.2)

        t = threading.Thread(target=loop, daemon=True)
        t.start()
        time.sleep(0.5)
        print("NVML thread     ", steps(ctx, 12), "samples", len(rows), rows[-1], flush=True)
        stop = True
    except Exception as e:  # noqa: BLE001
        print("pynvml failed", repr(e))
    print("no sampler again", steps(ctx, 12), flush=True)
