"""Does the step time depend on the ORDER of the points?  (random, as bench.py; sorted along x; sorted by 0.25 m voxel;
sorted by 2 m tile then random inside: a tiled TLS export.)"""
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
s, r, l, u, i = synth.cylinder_arrays(qsm)
eng.set_cylinders(*(torch.tensor(x, device=dev) for x in (s, r, l, u)), torch.tensor(i, device=dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
vox = np.floor((base - base.min(0)) / 0.25).astype(np.int64)
tile = np.floor((base - base.min(0)) / 2.0).astype(np.int64)
orders = {"random": np.arange(n), "sorted_x": np.argsort(base[:, 0], kind="stable"),
          "sorted_voxel": np.lexsort((vox[:, 0], vox[:, 1], vox[:, 2])), "tiles_2m": np.lexsort((tile[:, 0], tile[:, 1], tile[:, 2]))}
for name, order in orders.items():
    d = torch.tensor(base[order], device=dev)
    out = {"index": torch.empty(n, dtype=torch.int32, device=dev), "id": torch.empty(n, dtype=torch.int32, device=dev),
           "dist": torch.empty(n, dtype=torch.float32, device=dev), "offset": torch.empty((n, 3), dtype=torch.float32, device=dev)}
    for _ in range(2):
        eng.label(d, api.VARIANT_A, mode="grid", out=out, want=("index", "id", "dist", "offset"))
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.label(d, api.VARIANT_A, mode="grid", out=out, want=("index", "id", "dist", "offset")); b.record()
        torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    eng.set_profiling(True); eng.label(d, api.VARIANT_A, mode="grid", out=out, want=("index", "id", "dist", "offset")); ph = eng.phase_ms(); eng.set_profiling(False)
    print(json.dumps({"order": name, "ms": float(np.median(ms)), **{k: round(v, 4) for k, v in ph.items()}}), flush=True)
