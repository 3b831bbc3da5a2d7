#!/usr/bin/env python
"""Per-phase device time of one labelling call as the cloud shrinks: what a rank of a strongly-scaled 10M-point plot sees
(the fixed cost per call caps strong scaling).  Same table (50k cylinders), rows [0, n) of the bench
cloud, L2 flushed between calls.

    python scripts/bench_floor.py [--out gpurun_out/floor.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from treemorph_b200 import api, synth      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "floor.json"))
    ap.add_argument("--cylinders", type=int, default=50_000)
    ap.add_argument("--sizes", default="10000,100000,1000000,1250000,2500000,5000000,10000000")
    ap.add_argument("--cells", default="0", help="comma-separated voxel edges to try (0 = the library's own choice)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = api.Engine(dev)
    qsm = synth.random_qsm(args.cylinders, seed=1)
    sizes = [int(s) for s in args.sizes.split(",")]
    pts = torch.as_tensor(synth.sample_points(qsm, max(sizes), seed=2)).to(dev)
    start, radius, length, unit, ids = synth.cylinder_arrays(qsm)
    eng.set_cylinders(torch.tensor(start, device=dev), torch.tensor(radius, device=dev), torch.tensor(length, device=dev),
                      torch.tensor(unit, device=dev), torch.tensor(ids, device=dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for n, cell in [(n, float(c)) for n in sizes for c in args.cells.split(",")]:
        p = pts[:n]
        out = {"index": torch.empty(n, dtype=torch.int32, device=dev), "id": torch.empty(n, dtype=torch.int32, device=dev),
               "dist": torch.empty(n, dtype=torch.float32, device=dev), "offset": torch.empty((n, 3), dtype=torch.float32, device=dev)}
        for _ in range(3):
            eng.label(p, api.VARIANT_A, mode="grid", cell_size=cell, out=out)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        torch.cuda.synchronize()
        for a, b in evs:
            flush.fill_(1)
            a.record()
            eng.label(p, api.VARIANT_A, mode="grid", cell_size=cell, out=out)
            b.record()
        torch.cuda.synchronize()
        ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
        eng.set_profiling(True)
        acc = {}
        for _ in range(5):
            flush.fill_(1)
            eng.label(p, api.VARIANT_A, mode="grid", cell_size=cell, out=out)
            for k, v in eng.phase_ms().items():
                acc.setdefault(k, []).append(v)
        eng.set_profiling(False)
        row = {"points": n, "cylinders": args.cylinders, "cell": cell, "ms": ms, "points_per_s": n / (ms * 1e-3),
               "phases_ms": {k: round(float(np.median(v)), 4) for k, v in acc.items()}}
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
