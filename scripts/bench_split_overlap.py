#!/usr/bin/env python
"""Does a large cloud gain from being labelled as K slices on K streams at once?  The sort phase of a call is bound by L2
atomics, the tile kernels by instruction issue, the epilogue by HBM: slices in different phases could share the GPU.
K engines on one device (same table), slice k of the bench cloud on stream k, whole job timed with CUDA events around
the fork / join, L2 flushed between iterations.

    python scripts/bench_split_overlap.py [--points 10000000] [--ways 1,2,3,4]
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
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--cylinders", type=int, default=50_000)
    ap.add_argument("--ways", default="1,2,3,4")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "split_overlap.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    qsm = synth.random_qsm(args.cylinders, seed=1)
    n = args.points
    pts = torch.as_tensor(synth.sample_points(qsm, n, seed=2)).to(dev)
    start, radius, length, unit, ids = (torch.tensor(x, device=dev) for x in synth.cylinder_arrays(qsm))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ways = [int(w) for w in args.ways.split(",")]
    engines = []
    for _ in range(max(ways)):
        e = api.Engine(dev)
        e.set_cylinders(start, radius, length, unit, ids)
        engines.append(e)
    streams = [torch.cuda.Stream(dev) for _ in range(max(ways))]
    out = {"index": torch.empty(n, dtype=torch.int32, device=dev), "id": torch.empty(n, dtype=torch.int32, device=dev),
           "dist": torch.empty(n, dtype=torch.float32, device=dev), "offset": torch.empty((n, 3), dtype=torch.float32, device=dev)}
    engines[0].label(pts, api.VARIANT_A, mode="grid", out=out)
    want = out["index"].clone()
    rows = []
    for k in ways:
        bounds = [(n * j // k, n * (j + 1) // k) for j in range(k)]
        views = [{name: t[lo:hi] for name, t in out.items()} for lo, hi in bounds]

        def job(concurrent: bool):
            main_s = torch.cuda.current_stream(dev)
            if not concurrent:
                for j, (lo, hi) in enumerate(bounds):
                    engines[j].label(pts[lo:hi], api.VARIANT_A, mode="grid", out=views[j])
                return
            fork = torch.cuda.Event()
            fork.record(main_s)
            for j, (lo, hi) in enumerate(bounds):
                streams[j].wait_event(fork)
                with torch.cuda.stream(streams[j]):
                    engines[j].label(pts[lo:hi], api.VARIANT_A, mode="grid", out=views[j])
                done = torch.cuda.Event()
                done.record(streams[j])
                main_s.wait_event(done)

        for concurrent in ([False] if k == 1 else [False, True]):
            for _ in range(3):
                job(concurrent)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
            torch.cuda.synchronize()
            for a, b in evs:
                flush.fill_(1)
                a.record()
                job(concurrent)
                b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
            same = bool(torch.equal(out["index"], want))
            row = {"points": n, "slices": k, "concurrent": concurrent, "ms": ms, "points_per_s": n / (ms * 1e-3), "same_rows": same}
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
