#!/usr/bin/env python
"""Small labelling calls for a quick self-check (and compute-sanitizer where it is available): the sorted path, the direct path, a table
with special / axis-parallel cylinders, clutter and non-finite points.  Run as
    TM_DIRECT=0 compute-sanitizer --tool memcheck python scripts/selfcheck_small.py
    TM_DIRECT=1 compute-sanitizer --tool racecheck python scripts/selfcheck_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from treemorph_b200 import api, synth

dev = torch.device("cuda", 0)
eng = api.Engine(dev)
q = synth.random_qsm(1200, seed=3)
pts = synth.sample_points(q, 60_000, seed=4)
rng = np.random.default_rng(5)
pts[::11] += rng.normal(0, 1.5, size=pts[::11].shape).astype(np.float32)
pts[3] = [np.nan, 0, 0]
pts[4] = [np.inf, 1, 2]
s, r, l, u, i = synth.cylinder_arrays(q)
u[7] = [0, 0, 1]                      # axis-parallel
l[9] = 0.0; u[9] = [np.nan] * 3       # zero-length cylinder of variant A: special
for vn in "AB":
    if vn == "B":
        u[9] = [0, 0, 0]
    eng.set_cylinders(*[torch.tensor(x, device=dev) for x in (s, r, l, u)], torch.tensor(i, device=dev))
    g = eng.label(torch.tensor(pts, device=dev), api.VARIANTS[vn], mode="grid")
    b = eng.label(torch.tensor(pts[:8000], device=dev), api.VARIANTS[vn], mode="brute")
    torch.cuda.synchronize()
    same = bool((g["index"][:8000] == b["index"]).all())
    print(vn, "grid == brute on 8000 rows:", same, eng.stats()["points_slow"], flush=True)
    assert same
rec = eng.label_cloud_host(pts.astype(np.float64), api.VARIANT_B, tail=(1.0, 1.0, 1.0, 1.0))
assert rec.shape == (len(pts), 11)
print("ok")
