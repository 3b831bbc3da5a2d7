"""Latency of the small-table caller (cylinder_proximity_based_segmentation pattern: 5 cylinders against rows of a resident
1M-point cloud): the kernel reading / writing page-locked host memory itself (default up to 64k rows) versus staged DMA copies
(TM_SMALL_STAGED=1)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from treemorph_b200 import api, synth
eng = api.get_engine(torch.device("cuda", 0))
qsm = synth.random_qsm(5000, seed=7)
cloud = synth.sample_points(qsm, 1_000_000, seed=8).astype(np.float64)
eng.upload_cloud(cloud)
start = np.stack([qsm["startX"], qsm["startY"], qsm["startZ"]], 1)[100:105].astype(np.float32)
end = np.stack([qsm["endX"], qsm["endY"], qsm["endZ"]], 1)[100:105].astype(np.float32)
radius = np.asarray(qsm["radius"])[100:105].astype(np.float32)
rng = np.random.default_rng(1)
ref = {}
for staged in ("1", None):
    if staged:
        os.environ["TM_SMALL_STAGED"] = staged
    else:
        os.environ.pop("TM_SMALL_STAGED", None)
    for n in (200, 2_000, 20_000, 50_000):
        rows = np.sort(np.random.default_rng(n).choice(len(cloud), n, replace=False))
        for _ in range(20):
            out = eng.proximity_flags(rows, start, end, radius, 0.05, want_dist=True, want_index=True)
        key = n
        if key in ref:
            same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out, ref[key]))
        else:
            ref[key] = out
            same = None
        reps = 500
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.proximity_flags(rows, start, end, radius, 0.05)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"path": "staged copies" if staged else "kernel reads/writes pinned host memory", "rows": n,
                          "us_per_call": round(dt * 1e6, 1), "same_as_staged": same}), flush=True)
