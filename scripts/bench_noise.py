"""Device generation of noisy surface clouds (tm_noise_cloud) versus the reference's numpy formulation on the host cores.

    python scripts/bench_noise.py [--points 10000000 100000000]

A 50k-cylinder plot; the per-cylinder counts of the reference are scaled so that the cloud has the requested size.
Device time with CUDA events (median of 7 after 3 warm-ups; the outputs, 24 + 12 B/point, are larger than L2 from 4M points up).
The host figure runs the numpy restatement of the reference (oracle/noise_cloud.py: np.repeat, three np.random draws, fancy-
indexed (N,3,3) rotations, einsum) on a bounded sample.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from treemorph_b200 import api, synth
from treemorph_b200.PreProcessing import NoiseDataGeneration as N
from oracle import noise_cloud as nc

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, nargs="+", default=[1_000_000, 10_000_000, 100_000_000])
ap.add_argument("--out", default="gpurun_out/noise.json")
args = ap.parse_args()
dev = torch.device("cuda", 0)
eng = api.get_engine(dev)
df = synth.qsm_dataframe(synth.random_qsm(50_000, seed=1))
t0 = time.perf_counter()
base = N.cylinder_plan(df)
plan_ms = 1e3 * (time.perf_counter() - t0)
res = {"cylinders": len(df), "host_plan_ms": plan_ms, "reference_density_points": base.n_points, "runs": []}
for n_target in args.points:
    hp = N.CylinderPlan(base.records, np.maximum(1, (base.counts * (n_target / base.n_points)).astype(np.int64)), base.first_point.copy())
    hp.first_point[1:] = np.cumsum(hp.counts)
    n = hp.n_points
    rec, first = torch.from_numpy(hp.records).to(dev), torch.from_numpy(hp.first_point).to(dev)
    for want32 in (False, True):
        ms = []
        for rep in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            out = eng.noise_cloud(rec, first, n=n, seed=rep, want_f32=want32)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            del out
        t = float(np.median(ms[3:]))
        bytes_pt = 24 + (12 if want32 else 0)
        res["runs"].append({"points": n, "f32_copy": want32, "device_ms": t, "points_per_s": n / (t * 1e-3),
                            "write_GB_per_s": n * bytes_pt / (t * 1e-3) / 1e9})
        print(json.dumps(res["runs"][-1]), flush=True)
# host: the reference's formulation on a bounded sample
op = nc.plan(df[["startX", "startY", "startZ"]].values, df[["endX", "endY", "endZ"]].values, df["radius"].values)
op.count = np.maximum(1, (op.count * (2_000_000 / op.count.sum())).astype(np.int64))
t0 = time.perf_counter()
np.random.seed(1)
cloud = nc.place(op, *nc.legacy_variates(op))
dt = time.perf_counter() - t0
res["host_numpy"] = {"points": int(len(cloud)), "s": dt, "points_per_s": len(cloud) / dt, "threads": 1}
print(json.dumps(res["host_numpy"]), flush=True)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(res, open(args.out, "w"), indent=1)
