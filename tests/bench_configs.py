#!/usr/bin/env python
"""BASELINE.json configs[1], [3], [4] and the small-table caller on ONE B200 (bench.py measures configs[2]).

    python tests/bench_configs.py --out gpurun_out/configs.json [--max-points 100000000] [--skip-sweep]

config 1 : single tree, 100k points x 2k cylinders (the reference's own CPU-runnable case): device resident, through the
           LabelGenerationCuda drop-in, and the oracle port on the host cores; all 100k rows compared bit-for-bit.
config 2 : 1M points x 10k cylinders, variants A and B, labels + offsets, device resident; checked against the
           oracle on a subsample.
config 4 : 5M PTv3-style corrected points x 50k cylinders through the Projection drop-in
           (generate_offset_cloud_cuda_batched, host cloud in, (N,7) float64 out); checked against the oracle on a
           20k-point subsample.
config 5 : sweep N in {1e5, 1e6, 1e7, 1e8} x M in {1e3, 1e4, 5e4, 2e5}, device resident, auto mode; every cell checked
           against the exhaustive GPU kernel on a subsample (the exhaustive kernel itself is pinned to the oracle by
           the test-suite) and against the oracle on a smaller one.
small    : cylinder_proximity_based_segmentation call pattern (M = 5, n = 2k / 50k rows of a resident 1M cloud):
           calls per second through tm_proximity_flags_host.

Lives under tests/ because it uses the CPU oracle as its checker (only tests/, smoke() and bench.py's CPU legs may).

Timing: CUDA events, 3 warm-up + 5 timed repetitions, L2 flushed (256 MiB write) between repetitions; host paths by
wall clock around synchronous calls.  Not part of the driver's contract; results are committed under profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, fn, flush, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms))


def install(eng, torch, arrs):
    start, radius, length, unit, ids = arrs
    dev = eng.device
    eng.set_cylinders(torch.tensor(start, device=dev), torch.tensor(radius, device=dev), torch.tensor(length, device=dev),
                      torch.tensor(unit, device=dev), torch.tensor(ids, device=dev))


def check_subsample(eng, torch, api, oracle, var, ovar, arrs, pts, dpts, full, n_gpu=20_000, n_cpu=2_000, seed=3):
    rng = np.random.default_rng(seed)
    n = len(pts)
    sub = rng.choice(n, min(n, n_gpu), replace=False)
    dsub = torch.tensor(sub, device=eng.device)
    brute = eng.label(dpts[dsub], var, mode="brute", want=("index", "dist", "offset"))
    ok_gpu = all(bool(((full[k][dsub] == brute[k]) | (full[k][dsub] != full[k][dsub])).all()) for k in ("index", "dist", "offset"))
    osub = sub[:n_cpu]
    with np.errstate(all="ignore"):
        ora = oracle.label(pts[osub], *arrs, ovar)
    got = {k: full[k][dsub[:n_cpu]].cpu().numpy() for k in ("index", "dist", "offset")}
    ok_cpu = bool((got["index"] == ora["index"]).all() and np.array_equal(got["dist"], ora["dist"], equal_nan=True)
                  and np.array_equal(got["offset"], ora["offset"], equal_nan=True))
    return {"vs_exhaustive_gpu": ok_gpu, "rows_gpu": int(len(sub)), "vs_oracle_bitwise": ok_cpu, "rows_oracle": int(len(osub))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--max-points", type=int, default=100_000_000)
    ap.add_argument("--skip-sweep", action="store_true")
    args = ap.parse_args()
    import torch
    from oracle import oracle
    from treemorph_b200 import api, synth
    from treemorph_b200.Modules import Projection
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    eng = api.Engine(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"gpu": torch.cuda.get_device_name(0), "host_cpus": os.cpu_count()}
    want = ("index", "id", "dist", "offset")

    # ---- config 1: single tree, 100k points x 2k cylinders, label generation; EVERY point checked against the oracle
    qsm = synth.random_qsm(2_000, seed=1)
    pts = synth.sample_points(qsm, 100_000, seed=2)
    arrs = synth.cylinder_arrays(qsm)
    install(eng, torch, arrs)
    dpts = torch.tensor(pts, device=dev)
    med, best = timed(torch, lambda: eng.label(dpts, api.VARIANT_A, mode="auto", want=want), flush)
    full = {k: v.cpu().numpy() for k, v in eng.label(dpts, api.VARIANT_A, mode="auto", want=want).items()}
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        ora = oracle.label(pts, *arrs, oracle.VARIANT_A)
    cpu_s = time.perf_counter() - t0
    from treemorph_b200.PreProcessing import LabelGenerationCuda
    df = synth.qsm_dataframe(qsm)
    cloud64 = pts.astype(np.float64)
    LabelGenerationCuda.generate_offset_cloud_cuda_batched(cloud64, df, dev)
    api_times = []
    for _ in range(7):
        t0 = time.perf_counter()
        rec = LabelGenerationCuda.generate_offset_cloud_cuda_batched(cloud64, df, dev)
        api_times.append(time.perf_counter() - t0)
    api_s = float(np.median(api_times))
    res["config1"] = {"points": 100_000, "cylinders": 2_000, "device_resident_ms": med, "points_per_s": 1e5 / (med * 1e-3),
                      "dropin_generate_offset_cloud_ms": api_s * 1e3,
                      "oracle_port_cpu_s": cpu_s, "oracle_threads": oracle.max_threads(),
                      "reference_cpu_s_in_build_container": "~165 s on 8 threads (SURVEY.md B.4)",
                      "all_rows_bitwise_equal_to_oracle": bool((full["index"] == ora["index"]).all() and np.array_equal(full["dist"], ora["dist"], equal_nan=True)
                                                               and np.array_equal(full["offset"], ora["offset"], equal_nan=True)
                                                               and np.array_equal(rec[:, 3:6], ora["offset"].astype(np.float64), equal_nan=True)
                                                               and np.array_equal(rec[:, 6], ora["id"].astype(np.float64)))}
    print("config1", json.dumps(res["config1"]), flush=True)

    # ---- config 2
    qsm = synth.random_qsm(10_000, seed=1)
    pts = synth.sample_points(qsm, 1_000_000, seed=2)
    dpts = torch.tensor(pts, device=dev)
    res["config2"] = {}
    for vn in "AB":
        var, ovar = api.VARIANTS[vn], oracle.VARIANTS[vn]
        arrs = synth.cylinder_arrays(qsm, ovar.axis_eps)
        install(eng, torch, arrs)
        eng.label(dpts[:4096], var, mode="grid", want=("id",))
        med, best = timed(torch, lambda: eng.label(dpts, var, mode="grid", want=want), flush)
        full = eng.label(dpts, var, mode="grid", want=want)
        st = eng.stats()
        res["config2"][vn] = {"points": 1_000_000, "cylinders": 10_000, "ms_median": med, "ms_min": best,
                              "points_per_s": 1e6 / (med * 1e-3), "pairs_evaluated": st["pairs_evaluated"],
                              "check": check_subsample(eng, torch, api, oracle, var, ovar, arrs, pts, dpts, full)}
    print("config2", json.dumps(res["config2"]), flush=True)

    # ---- config 4 (Projection drop-in, host in / host out)
    qsm = synth.random_qsm(50_000, seed=1)
    pts = synth.sample_points(qsm, 5_000_000, seed=4, noise="model")
    df = synth.qsm_dataframe(qsm)
    cloud64 = pts.astype(np.float64)
    t0 = time.perf_counter()
    rec = Projection.generate_offset_cloud_cuda_batched(cloud64, df, dev)        # first call: allocations, index build
    cold = time.perf_counter() - t0
    warm = []
    for _ in range(3):
        t0 = time.perf_counter()
        rec = Projection.generate_offset_cloud_cuda_batched(cloud64, df, dev)
        warm.append(time.perf_counter() - t0)
    dt = float(np.median(warm))
    rng = np.random.default_rng(9)
    sub = rng.choice(len(pts), 20_000, replace=False)
    arrs = synth.cylinder_arrays(qsm, oracle.VARIANT_B.axis_eps)
    with np.errstate(all="ignore"):
        ora = oracle.label(pts[sub], *arrs, oracle.VARIANT_B)
    ok = bool(np.array_equal(rec[sub, 6], ora["id"].astype(np.float64)) and np.array_equal(rec[sub, 3:6], ora["offset"].astype(np.float64), equal_nan=True)
              and np.array_equal(rec[sub, :3], cloud64[sub]))
    res["config4"] = {"points": 5_000_000, "cylinders": 50_000, "api": "Modules.Projection.generate_offset_cloud_cuda_batched (float64 host cloud -> (N,7) float64)",
                      "seconds": dt, "seconds_first_call": cold, "points_per_s": 5e6 / dt, "vs_oracle_bitwise_rows": 20_000, "vs_oracle_bitwise": ok}
    print("config4", json.dumps(res["config4"]), flush=True)
    del rec, cloud64

    # ---- small-table caller
    qsm = synth.random_qsm(5000, seed=7)
    cloud = synth.sample_points(qsm, 1_000_000, seed=8).astype(np.float64)
    t0 = time.perf_counter()
    eng.upload_cloud(cloud)
    up_s = time.perf_counter() - t0
    start = np.stack([qsm["startX"], qsm["startY"], qsm["startZ"]], 1)[100:105].astype(np.float32)
    end = np.stack([qsm["endX"], qsm["endY"], qsm["endZ"]], 1)[100:105].astype(np.float32)
    radius = np.asarray(qsm["radius"])[100:105].astype(np.float32)
    res["small_table"] = {"resident_cloud_rows": 1_000_000, "upload_s": up_s, "cylinders": 5}
    rng = np.random.default_rng(1)
    for n in (2_000, 50_000):
        rows = np.sort(rng.choice(len(cloud), n, replace=False))
        for _ in range(20):
            eng.proximity_flags(rows, start, end, radius, 0.05)
        reps = 300
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.proximity_flags(rows, start, end, radius, 0.05)
        dt = (time.perf_counter() - t0) / reps
        res["small_table"][f"rows_{n}"] = {"us_per_call": dt * 1e6, "calls_per_s": 1.0 / dt, "points_per_s": n / dt}
    print("small", json.dumps(res["small_table"]), flush=True)

    # ---- feature step (Modules/Features.add_features as the drivers call it: normals k=15 + relative height)
    from treemorph_b200.Modules import Features
    qsm = synth.random_qsm(5000, seed=7)
    res["features"] = {}
    for n in (1_000_000, 10_000_000):
        cloud = synth.sample_points(qsm, n, seed=8).astype(np.float64)
        dp = torch.tensor(cloud, device=dev)
        eng.knn_covariance(dp, 15)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.knn_covariance(dp, 15); b.record(); torch.cuda.synchronize()
        knn_ms = a.elapsed_time(b)
        a.record(); eng.radius_count(dp, 0.1); b.record(); torch.cuda.synchronize()
        rc_ms = a.elapsed_time(b)
        entry = {"knn15_covariance_device_ms": knn_ms, "radius_count_device_ms": rc_ms, "points_per_s_knn": n / (knn_ms * 1e-3)}
        if n == 1_000_000:
            lab = np.concatenate([cloud, np.zeros((n, 4))], axis=1)
            t0 = time.perf_counter()
            out = Features.add_features(lab, use_densities=False, use_curvatures=False, use_distances=False, use_verticalities=False)
            entry["add_features_normals_height_s"] = time.perf_counter() - t0
            entry["add_features_shape"] = list(out.shape)
            from scipy.spatial import cKDTree
            t0 = time.perf_counter()
            tree = cKDTree(cloud[:100_000])
            tree.query(cloud[:100_000], k=15)
            entry["scipy_ckdtree_build_query_100k_s"] = time.perf_counter() - t0
        res["features"][str(n)] = entry
        del dp
    print("features", json.dumps(res["features"]), flush=True)

    # ---- config 5 sweep
    if not args.skip_sweep:
        cells = []
        for ci, m in enumerate((1_000, 10_000, 50_000, 200_000)):
            qsm = synth.random_qsm(m, seed=1000 + ci)
            arrs = synth.cylinder_arrays(qsm)
            install(eng, torch, arrs)
            n_max = min(args.max_points, 100_000_000)
            pts_all = synth.sample_points(qsm, n_max, seed=2000 + ci)
            dall = torch.tensor(pts_all, device=dev)
            for n in (100_000, 1_000_000, 10_000_000, 100_000_000):
                if n > n_max:
                    continue
                dpts, pts = dall[:n], pts_all[:n]
                out = {k: torch.empty(s, dtype=t, device=dev) for k, (s, t) in
                       {"index": ((n,), torch.int32), "id": ((n,), torch.int32), "dist": ((n,), torch.float32),
                        "offset": ((n, 3), torch.float32)}.items()}
                med, best = timed(torch, lambda: eng.label(dpts, api.VARIANT_A, mode="auto", want=want, out=out), flush,
                                  reps=3 if n >= 100_000_000 else 5, warm=2)
                st = eng.stats()
                chk = check_subsample(eng, torch, api, oracle, api.VARIANT_A, oracle.VARIANT_A, arrs, pts, dpts, out,
                                      n_gpu=20_000, n_cpu=1_000 if m >= 50_000 else 2_000)
                cell = {"points": n, "cylinders": m, "mode_used": {1: "brute", 2: "grid"}[st["mode_used"]], "ms_median": med,
                        "ms_min": best, "points_per_s": n / (med * 1e-3), "pairs_evaluated": st["pairs_evaluated"],
                        "cell_size": st["cell_size"], "brute_equiv_pairs_per_s": n * m / (med * 1e-3), "points_ring": st["points_ring"], "points_tree": st["points_tree"],
                        "points_brute": st["points_brute"], "check": chk}
                cells.append(cell)
                print("sweep", json.dumps(cell), flush=True)
                del out
            del dall, pts_all
            torch.cuda.empty_cache()
        res["config5_sweep"] = cells

    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", args.out)
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
