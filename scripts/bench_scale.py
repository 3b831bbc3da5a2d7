"""The same 1M x 10k workload in different UNITS (model scaled by 0.001 / 0.01 / 1 / 100): the automatic voxel edge follows the
cylinder sizes, so the step time must not depend on the unit.  Each run is checked against the exhaustive kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from treemorph_b200 import api, synth
dev = torch.device("cuda", 0)
eng = api.Engine(dev)
qsm = synth.random_qsm(10_000, seed=1)
pts0 = synth.sample_points(qsm, 1_000_000, seed=2).astype(np.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for scale in (0.001, 0.01, 1.0, 100.0):
    q = {k: (np.asarray(v) * scale if k != "ID" else v) for k, v in qsm.items()}
    s, r, l, u, i = synth.cylinder_arrays(q)
    eng.set_cylinders(*(torch.tensor(x, device=dev) for x in (s, r, l, u)), torch.tensor(i, device=dev))
    d = torch.tensor((pts0 * scale).astype(np.float32), device=dev)
    for _ in range(2):
        g = eng.label(d, api.VARIANT_A, mode="grid", want=("index", "dist"))
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g = eng.label(d, api.VARIANT_A, mode="grid", want=("index", "dist")); b.record()
        torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    st = eng.stats()
    sub = torch.arange(0, len(d), 50, device=dev)
    bb = eng.label(d[sub], api.VARIANT_A, mode="brute", want=("index", "dist"))
    ok = bool((bb["index"] == g["index"][sub]).all()) and bool(((bb["dist"] == g["dist"][sub]) | torch.isnan(bb["dist"])).all())
    print(json.dumps({"scale": scale, "ms": float(np.median(ms)), "cell_size": st["cell_size"], "pairs_per_point": st["pairs_evaluated"] / len(d),
                      "culls_per_point": st["cull_tests"] / len(d), "equals_exhaustive": ok,
                      "points_grid": st["points_grid"], "points_far": st["points_far"], "points_ring": st["points_ring"],
                      "points_tree": st["points_tree"], "index_entries": st["index_entries"], "voxels_occupied": st["voxels_occupied"],
                      "grid_dim": st["grid_dim"]}), flush=True)
