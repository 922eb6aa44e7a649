"""Micro-benchmarks behind the FP64 roofline: pipe peaks, instruction costs (prints JSON)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nmfk.jl_b200", "python"))
import nmfk_b200 as nb  # noqa: E402

names = {0: "dfma_tflops", 1: "dmma_tflops", 2: "ffma_tflops", 3: "copy_gbs", 4: "dfma_latency_cyc", 5: "ffma_latency_cyc",
         6: "rcp64h_plus_dfma_latency_cyc", 7: "rcp64h_cyc_per_warp_instr_per_smsp", 8: "rcp32_cyc_per_warp_instr_per_smsp",
         9: "dmma_plus_dfma_mixed_tflops", 10: "kl_reciprocal_seq_cyc_per_warp_per_smsp"}
with nb.Context(0) as ctx:
    out = {}
    for w, nm in names.items():
        out[nm] = round(max(ctx.measure_peak(w) for _ in range(2)) if w in (0, 1, 2, 3, 9) else min(ctx.measure_peak(w) for _ in range(2)), 3)
    print(json.dumps(out))
