"""How does the step time degrade when a fraction of the cloud is far from every cylinder (foliage, ground, clutter)?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from treemorph_b200 import api, synth
dev = torch.device("cuda", 0)
eng = api.Engine(dev)
qsm = synth.random_qsm(50_000, seed=1)
n = 10_000_000
base = synth.sample_points(qsm, n, seed=2)
lo, hi = base.min(0), base.max(0)
s, r, l, u, i = synth.cylinder_arrays(qsm)
eng.set_cylinders(*(torch.tensor(x, device=dev) for x in (s, r, l, u)), torch.tensor(i, device=dev))
rng = np.random.default_rng(5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for frac in (0.0, 0.01, 0.05, 0.2, 0.5):
    pts = base.copy()
    k = int(frac * n)
    if k:
        rows = rng.choice(n, k, replace=False)
        pts[rows] = (lo + rng.random((k, 3)) * (hi - lo)).astype(np.float32)       # uniform in the plot's bounding box
    d = torch.tensor(pts, device=dev)
    out = {"index": torch.empty(n, dtype=torch.int32, device=dev), "id": torch.empty(n, dtype=torch.int32, device=dev),
           "dist": torch.empty(n, dtype=torch.float32, device=dev), "offset": torch.empty((n, 3), dtype=torch.float32, device=dev)}
    for _ in range(2):
        eng.label(d, api.VARIANT_A, mode="grid", out=out, want=("index", "id", "dist", "offset"))
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.label(d, api.VARIANT_A, mode="grid", out=out, want=("index", "id", "dist", "offset")); b.record()
        torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    st = eng.stats()
    eng.set_profiling(True); eng.label(d, api.VARIANT_A, mode="grid", out=out, want=("index", "id", "dist", "offset")); ph = eng.phase_ms(); eng.set_profiling(False)
    print(json.dumps({"clutter_fraction": frac, "ms": float(np.median(ms)), "points_far": st["points_far"], "points_ring": st["points_ring"], "points_tree": st["points_tree"],
                      "points_brute": st["points_brute"], "pairs": st["pairs_evaluated"], "culls": st["cull_tests"],
                      "evaluate_ms": ph["evaluate"], "tree_ms": ph["tree"], "exhaustive_ms": ph["exhaustive"]}), flush=True)
