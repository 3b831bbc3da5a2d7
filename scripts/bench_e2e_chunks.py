#!/usr/bin/env python
"""The drop-in call (pageable float64 cloud, DataFrame) as a function of the host pipeline's chunk size and depth
(TM_HOST_CHUNK / TM_HOST_DEPTH are read by the library on every call), with the library's own lap times (TM_TRACE_HOST).

    python scripts/bench_e2e_chunks.py [--points 10000000]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from treemorph_b200 import synth      # noqa: E402
from treemorph_b200.PreProcessing import LabelGenerationCuda as L      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--cylinders", type=int, default=50_000)
    ap.add_argument("--chunks", default="262144,524288,1048576,2097152")
    ap.add_argument("--depths", default="4")
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e_chunks.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    qsm = synth.random_qsm(args.cylinders, seed=1)
    cloud = np.ascontiguousarray(synth.sample_points(qsm, args.points, seed=2), dtype=np.float64)
    df = synth.qsm_dataframe(qsm)
    rows = []
    settings = [(c, d) for d in args.depths.split(",") for c in args.chunks.split(",")]
    times = {k: [] for k in settings}
    for rep in range(args.reps + 2):                       # interleaved: every setting sees the same drift of the box
        for chunk, depth in settings:
            os.environ["TM_HOST_CHUNK"], os.environ["TM_HOST_DEPTH"] = chunk, depth
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            L.generate_offset_cloud_cuda_batched(cloud, df, dev)
            torch.cuda.synchronize()
            if rep >= 2:
                times[(chunk, depth)].append(time.perf_counter() - t0)
    for (chunk, depth), t in times.items():
        row = {"chunk": int(chunk), "depth": int(depth), "ms_median": 1e3 * float(np.median(t)), "ms_min": 1e3 * min(t),
               "points_per_s": args.points / float(np.median(t))}
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.environ["TM_TRACE_HOST"] = "1"
    os.environ["TM_HOST_CHUNK"], os.environ["TM_HOST_DEPTH"] = "1048576", "4"
    L.generate_offset_cloud_cuda_batched(cloud, df, dev)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
