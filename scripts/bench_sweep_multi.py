#!/usr/bin/env python
"""BASELINE.json configs[4] across GPUs: one cloud of N points against M cylinders, the rows sharded over the ranks of
one node (strong scaling), device resident, L2 flushed between steps, CUDA events, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
        scripts/bench_sweep_multi.py --out gpurun_out/sweep_wW.json

Rank 0 reads the QSM and the table travels through tm_broadcast_cylinders (C ABI, NCCL); every rank samples its own
rows of the plot (same distribution, its own seed) and checks 20 000 of them bit for bit against the exhaustive kernel.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from treemorph_b200 import api, sharding, synth      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_multi.json"))
    ap.add_argument("--points", default="1000000,10000000,100000000")
    ap.add_argument("--cylinders", default="10000,50000")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    eng = api.Engine(dev)
    sharding.init_engine_comm(eng)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cells = []
    want = ("index", "id", "dist", "offset")
    for ci, m in enumerate(int(v) for v in args.cylinders.split(",")):
        qsm = synth.random_qsm(m, seed=1000 + ci)
        if rank == 0:
            start, radius, length, unit, ids = synth.cylinder_arrays(qsm)
            eng.broadcast_cylinders(torch.tensor(start, device=dev), torch.tensor(radius, device=dev), torch.tensor(length, device=dev),
                                    torch.tensor(unit, device=dev), torch.tensor(ids, device=dev), root=0)
        else:
            eng.broadcast_cylinders(root=0)
        totals = [int(v) for v in args.points.split(",")]
        n_max = max(sharding.shard_bounds(t, world, rank)[1] - sharding.shard_bounds(t, world, rank)[0] for t in totals)
        mine_all = torch.tensor(synth.sample_points(qsm, n_max, seed=2000 + ci + 101 * rank), device=dev)
        for total in totals:
            lo, hi = sharding.shard_bounds(total, world, rank)
            n = hi - lo
            pts = mine_all[:n]
            out = {"index": torch.empty(n, dtype=torch.int32, device=dev), "id": torch.empty(n, dtype=torch.int32, device=dev),
                   "dist": torch.empty(n, dtype=torch.float32, device=dev), "offset": torch.empty((n, 3), dtype=torch.float32, device=dev)}
            for _ in range(3):
                flush.fill_(1)
                eng.label(pts, api.VARIANT_A, mode="auto", want=want, out=out)
            reps = 5
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            for a, b in evs:
                flush.fill_(1)
                a.record()
                eng.label(pts, api.VARIANT_A, mode="auto", want=want, out=out)
                b.record()
            torch.cuda.synchronize()
            t = torch.tensor([float(np.median([a.elapsed_time(b) for a, b in evs]))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            # parity on this rank's rows: 20 000 of them against the exhaustive kernel, bit for bit
            rng = np.random.default_rng(5 + rank)
            sub = torch.tensor(rng.choice(n, min(n, 20_000), replace=False), device=dev)
            ref = eng.label(pts[sub], api.VARIANT_A, mode="brute", want=want)
            same = all(torch.equal(ref[k].view(torch.int32), out[k][sub].view(torch.int32)) for k in want)
            ok = torch.tensor([1 if same else 0], device=dev)
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if rank == 0:
                cell = {"gpus": world, "points_total": total, "points_per_rank": n, "cylinders": m, "ms": ms,
                        "points_per_s": total / (ms * 1e-3), "rows_checked_bitwise_on_every_rank": int(len(sub)), "all_ranks_equal_exhaustive": bool(ok.item())}
                cells.append(cell)
                print(json.dumps(cell), flush=True)
            del out
        del mine_all
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(cells, f, indent=1)
    eng.comm_destroy()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
