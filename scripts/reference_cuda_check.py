#!/usr/bin/env python
"""What does the reference compute on a B200?  

Runs ``oracle/torch_mirror.py`` — the reference's ATen call chain, op for op — on ``cuda`` and on ``cpu`` with the
DataFrame (Fortran-ordered) table layout, and diffs both against the C oracle (both ``norm`` roundings) and against this
repo's CUDA kernels; then times the torch-CUDA chain at the BASELINE configs next to this repo's path.  Writes one JSON
document (``gpurun_out/reference_cuda.json`` by default; the committed copy lives in ``profiles/``).

    python scripts/reference_cuda_check.py [--out PATH] [--quick]
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

from oracle import oracle, torch_mirror          # noqa: E402  (test infrastructure: this script is a checker)
from treemorph_b200 import api, synth           # noqa: E402


def diff(a: dict, b: dict) -> dict:
    idx = a["index"] != b["index"]
    both = ~idx
    da, db = a["dist"], b["dist"]
    same_d = (da == db) | (np.isnan(da) & np.isnan(db))
    oa, ob = a["offset"], b["offset"]
    same_o = (oa == ob) | (np.isnan(oa) & np.isnan(ob))
    with np.errstate(invalid="ignore"):
        dmax = float(np.nanmax(np.abs(da[both].astype(np.float64) - db[both].astype(np.float64)))) if both.any() else 0.0
        omax = float(np.nanmax(np.abs(oa[both].astype(np.float64) - ob[both].astype(np.float64)))) if both.any() else 0.0
    return {"n": int(len(da)), "index_mismatches": int(idx.sum()), "dist_not_bitwise": int((~same_d).sum()),
            "offset_rows_not_bitwise": int((~same_o.all(axis=1)).sum()), "max_abs_dist_diff_m": dmax,
            "max_abs_offset_diff_m": omax, "nan_pattern_equal": bool((np.isnan(da) == np.isnan(db)).all())}


def ours(eng, qsm, pts, var, norm_fma, ovar):
    dev = eng.device
    start, radius, length, unit, ids = synth.cylinder_arrays(qsm, ovar.axis_eps)
    eng.set_cylinders(torch.tensor(start, device=dev), torch.tensor(radius, device=dev), torch.tensor(length, device=dev),
                      torch.tensor(unit, device=dev), torch.tensor(ids, device=dev))
    got = eng.label(torch.tensor(pts, device=dev), var, norm_fma=norm_fma, mode="grid", want=("index", "id", "dist", "offset"))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in got.items()}


def parity_case(eng, m, n, seed, do_cpu_mirror=True):
    qsm = synth.random_qsm(m, seed=seed)
    pts = synth.sample_points(qsm, n, seed=seed + 1)
    cuda = torch.device("cuda", 0)
    res = {"points": n, "cylinders": m}
    for vn in "AB":
        ovar, var = oracle.VARIANTS[vn], api.VARIANTS[vn]
        arrs = synth.cylinder_arrays(qsm, ovar.axis_eps)
        mir_cuda_f = torch_mirror.label(pts, qsm, cuda, ovar.perp_atol, ovar.norm_eps, ovar.axis_eps, fortran=True)
        mir_cuda_c = torch_mirror.label(pts, qsm, cuda, ovar.perp_atol, ovar.norm_eps, ovar.axis_eps, fortran=False)
        ora0 = oracle.label(pts, *arrs, ovar, norm_fma=False)
        ora1 = oracle.label(pts, *arrs, ovar, norm_fma=True)
        our0 = ours(eng, qsm, pts, var, False, ovar)
        our1 = ours(eng, qsm, pts, var, True, ovar)
        # do the cylinder tensors built on the device equal the oracle's prep? (axis_length / axis_unit bits)
        st, rd, ln, un, _ = torch_mirror.table_tensors(qsm, cuda, ovar.axis_eps, fortran=True)
        prep = {"length_bitwise": bool(np.array_equal(ln.cpu().numpy().reshape(-1), np.asarray(arrs[2]).reshape(-1), equal_nan=True)),
                "unit_bitwise": bool(np.array_equal(un.cpu().numpy(), np.asarray(arrs[3]), equal_nan=True)),
                "start_stride_on_cuda": list(st.stride()), "unit_stride_on_cuda": list(un.stride())}
        entry = {"prep_vs_oracle": prep,
                 "torch_cuda_F_vs_oracle_norm_plain": diff(mir_cuda_f, ora0),
                 "torch_cuda_F_vs_oracle_norm_fma": diff(mir_cuda_f, ora1),
                 "torch_cuda_C_vs_oracle_norm_plain": diff(mir_cuda_c, ora0),
                 "torch_cuda_C_vs_oracle_norm_fma": diff(mir_cuda_c, ora1),
                 "torch_cuda_F_vs_torch_cuda_C": diff(mir_cuda_f, mir_cuda_c),
                 "ours_norm_plain_vs_torch_cuda_F": diff(our0, mir_cuda_f),
                 "ours_norm_fma_vs_torch_cuda_F": diff(our1, mir_cuda_f),
                 "ours_norm_plain_vs_oracle_norm_plain": diff(our0, ora0),
                 "ours_norm_fma_vs_oracle_norm_fma": diff(our1, ora1)}
        if do_cpu_mirror:
            mir_cpu_f = torch_mirror.label(pts, qsm, torch.device("cpu"), ovar.perp_atol, ovar.norm_eps, ovar.axis_eps, fortran=True)
            entry["torch_cpu_F_vs_oracle_norm_plain"] = diff(mir_cpu_f, ora0)
            entry["torch_cuda_F_vs_torch_cpu_F"] = diff(mir_cuda_f, mir_cpu_f)
        res[vn] = entry
    return res


def tie_case():
    """Duplicate cylinders: argmin must pick the lowest row on CUDA as on CPU."""
    qsm = synth.random_qsm(40, seed=3)
    dup = {k: np.concatenate([np.asarray(v), np.asarray(v)]) for k, v in qsm.items()}
    dup["ID"] = np.arange(80) + 7
    pts = synth.sample_points(qsm, 4000, seed=4)
    a = torch_mirror.label(pts, dup, torch.device("cuda", 0), 1e-6, 0.0, 0.0)
    return {"all_winners_in_first_copy": bool((a["index"] < 40).all())}


def time_reference_cuda(m, n_time, variant="A", fortran=True):
    qsm = synth.random_qsm(m, seed=1)
    pts = synth.sample_points(qsm, n_time, seed=2)
    ovar = oracle.VARIANTS[variant]
    cuda = torch.device("cuda", 0)
    torch_mirror.label(pts[:2048], qsm, cuda, ovar.perp_atol, ovar.norm_eps, ovar.axis_eps, fortran=fortran)      # warm-up
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    torch_mirror.label(pts, qsm, cuda, ovar.perp_atol, ovar.norm_eps, ovar.axis_eps, fortran=fortran)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"cylinders": m, "points_timed": n_time, "seconds": dt, "points_per_s": n_time / dt,
            "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "table_layout": "F" if fortran else "C", "variant": variant}


def time_ours(eng, m, n, variant="A"):
    """Same workload through this repo's drop-in call (pageable float64 host cloud in, (N,7) float64 out; table install
    included) and device resident."""
    import pandas as pd
    from treemorph_b200.PreProcessing import LabelGenerationCuda as dropA
    from treemorph_b200.Modules import Projection as dropB
    qsm = synth.random_qsm(m, seed=1)
    pts = synth.sample_points(qsm, n, seed=2).astype(np.float64)
    df = pd.DataFrame(qsm)
    mod = dropA if variant == "A" else dropB
    dev = torch.device("cuda", 0)
    mod.generate_offset_cloud_cuda_batched(pts, df, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        mod.generate_offset_cloud_cuda_batched(pts, df, dev)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    dpts = torch.tensor(pts, dtype=torch.float32, device=dev)
    var = api.VARIANTS[variant]
    e = api.get_engine(dev)
    e.label(dpts, var, mode="grid")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        e.label(dpts, var, mode="grid")
    e1.record()
    torch.cuda.synchronize()
    dms = e0.elapsed_time(e1) / 5
    return {"cylinders": m, "points": n, "dropin_seconds": dt, "dropin_points_per_s": n / dt,
            "device_resident_ms": dms, "device_resident_points_per_s": n / (dms * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_cuda.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    assert torch.cuda.is_available()
    oracle.build()
    eng = api.Engine(torch.device("cuda", 0))
    doc = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "host_cpus": os.cpu_count()}
    doc["ties"] = tie_case()
    doc["parity"] = [parity_case(eng, 2000, 20_000, 1)]
    if not args.quick:
        doc["parity"].append(parity_case(eng, 50_000, 40_000, 11, do_cpu_mirror=False))
    timing = []
    cases = [(2000, 100_000, 100_000), (10_000, 1_000_000, 200_000), (50_000, 10_000_000, 100_000), (50_000, 5_000_000, 100_000)]
    variants = ["A", "A", "A", "B"]
    if args.quick:
        cases, variants = cases[:1], variants[:1]
    for (m, n_cfg, n_time), vn in zip(cases, variants):
        ref = time_reference_cuda(m, n_time, vn)
        ref["config_points"] = n_cfg
        ref["extrapolated_seconds_at_config"] = ref["seconds"] * n_cfg / n_time
        mine = time_ours(eng, m, n_cfg, vn)
        timing.append({"reference_torch_cuda": ref, "this_repo": mine,
                       "speedup_dropin_call": ref["extrapolated_seconds_at_config"] / mine["dropin_seconds"],
                       "speedup_device_resident": ref["extrapolated_seconds_at_config"] / (mine["device_resident_ms"] * 1e-3)})
    doc["timing"] = timing
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc)[:4000])


if __name__ == "__main__":
    main()
