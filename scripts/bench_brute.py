#!/usr/bin/env python
"""The exhaustive kernel as FP32 yard-stick: pairs = N*M exactly, 81 lane-ops per pair, points per thread 1/2/4."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from treemorph_b200 import api, synth
dev = torch.device("cuda", 0)
eng = api.Engine(dev)
m = 50_000
qsm = synth.random_qsm(m, seed=1)
s, r, l, u, i = synth.cylinder_arrays(qsm)
eng.set_cylinders(*[torch.tensor(x, device=dev) for x in (s, r, l, u)], torch.tensor(i, device=dev))
peak = eng.fp32_peak()
for n in (400_000, 1_000_000):
    pts = torch.tensor(synth.sample_points(qsm, n, seed=2), device=dev)
    for p in (1, 2, 4):
        os.environ["TM_BRUTE_P"] = str(p)
        eng.label(pts, api.VARIANT_A, mode="brute", want=("id",))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.label(pts, api.VARIANT_A, mode="brute", want=("index", "id", "dist", "offset"))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"points": n, "cylinders": m, "points_per_thread": p, "ms": ms, "pairs_per_s": n * m / ms * 1e3,
                          "fp32_frac": n * m * 81 / (ms * 1e-3) / peak, "fp32_peak": peak}), flush=True)
