"""Generate the committed golden vectors by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the *inputs* (float64 QSM columns as ``pd.read_csv`` yields them, the cloud) and the
reference's *outputs* for both variants are stored:

    ref_A_cloud / ref_B_cloud : generate_offset_cloud_cuda_batched(cloud, df, cpu)   (N,7) float64
                                (LabelGenerationCuda.py:113-135, Projection.py:117-144)
    ref_A_id/dist/off, ref_B_*: closest_cylinder_cuda_batch(...) on tensors built exactly as
                                generate_offset_cloud_cuda_batched builds them (F-ordered start)

The reference has no tests or fixtures of its own for this path, hence this script.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "extracting-tree-morphology-from-point-clouds_b200"))

import ref_harness  # noqa: E402
import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(name: str, qsm: dict, cloud: np.ndarray) -> None:
    import pandas as pd
    import torch

    df = pd.DataFrame(qsm)
    rec = {"cloud": cloud}
    for k, v in qsm.items():
        rec["qsm_" + k] = np.asarray(v)
    for vn in "AB":
        with np.errstate(all="ignore"):
            rec[f"ref_{vn}_cloud"] = ref_harness.run_cloud(vn, cloud, df)
        # kernel level, tensors as the reference's own driver builds them
        start = torch.tensor(df[["startX", "startY", "startZ"]].values, dtype=torch.float32)
        end = torch.tensor(df[["endX", "endY", "endZ"]].values, dtype=torch.float32)
        radius = torch.tensor(df["radius"].values, dtype=torch.float32)
        ids = torch.tensor(df["ID"].values, dtype=torch.int32)
        axis = end - start
        length = torch.norm(axis, dim=1, keepdim=True)
        if vn == "A":
            unit = axis / length
        else:
            safe = length.clone()
            safe[safe < 1e-8] = 1e-8
            unit = axis / safe
        rid, rdist, roff = ref_harness.run_kernel(vn, cloud[:, :3], start, radius, length, unit, ids)
        rec[f"ref_{vn}_id"] = rid
        rec[f"ref_{vn}_dist"] = rdist
        rec[f"ref_{vn}_off"] = roff
        rec[f"ref_{vn}_length"] = length.numpy()
        rec[f"ref_{vn}_unit"] = unit.numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: N={len(cloud)} M={len(qsm['ID'])} -> {os.path.getsize(path)/1024:.0f} KiB")


def qsm_from_rows(rows, ids=None) -> dict:
    a = np.asarray(rows, dtype=np.float64)
    return {"startX": a[:, 0], "startY": a[:, 1], "startZ": a[:, 2], "endX": a[:, 3], "endY": a[:, 4],
            "endZ": a[:, 5], "radius": a[:, 6],
            "ID": np.asarray(ids if ids is not None else np.arange(len(a)), dtype=np.int64)}


def main() -> None:
    # 1. a small synthetic tree, NoiseDataGeneration-style cloud (float32 cloud)
    q = synth.random_qsm(300, seed=11)
    run_case("tree300", q, synth.sample_points(q, 1500, seed=12))

    # 2. same generator, IDs are arbitrary integers (offset + permuted), float64 cloud with extra columns
    q = synth.random_qsm(257, seed=21, id_offset=100)
    rng = np.random.default_rng(22)
    q["ID"] = rng.permutation(q["ID"])
    cloud = synth.sample_points(q, 1100, seed=23).astype(np.float64)
    cloud = np.concatenate([cloud + rng.normal(0, 1e-9, cloud.shape), rng.random((len(cloud), 2))], axis=1)
    run_case("tree257_ids_f64", q, cloud)

    # 3. two-tree plot, model-residual cloud (projection configuration)
    q = synth.random_qsm(400, seed=31)
    q2 = synth.random_qsm(300, seed=32, id_offset=400)
    for k in ("startX", "endX"):
        q2[k] = q2[k] + 4.0
    q = {k: np.concatenate([q[k], q2[k]]) for k in q}
    run_case("plot700_model", q, synth.sample_points(q, 1400, seed=33, noise="model"))

    # 4. adversarial: ties, on-axis points, caps, perp boundary, interior points
    rows = [
        [0, 0, 0, 0, 0, 1, 0.10],        # vertical stem
        [0, 0, 0, 0, 0, 1, 0.10],        # exact duplicate  → lowest row index must win
        [0, 0, 1, 0.5, 0, 1.5, 0.05],    # branch
        [2, 0, 0, 2, 0, 1, 0.30],        # thick stem (interior points)
        [2, 0, 0.2, 2.2, 0, 0.4, 0.02],  # twig starting inside the thick stem
        [5, 5, 5, 5.3, 5.4, 5.5, 0.01],  # far twig
    ]
    q = qsm_from_rows(rows, ids=[5, 6, 7, 8, 9, 10])
    pts = [
        [0.2, 0.0, 0.5],                  # plain mantle case, tie between rows 0/1
        [0.0, 0.0, 0.5],                  # exactly on the axis, inside the slab (A: NaN)
        [0.0, 0.0, 1.7],                  # on the axis line beyond the cap
        [0.05, 0.0, 1.2],                 # beyond cap, within radius  → cap disk
        [0.3, 0.1, 1.3],                  # beyond cap, outside radius → rim
        [0.05, 0.02, -0.4],               # below the start cap
        [2.05, 0.03, 0.5],                # interior of the thick stem
        [2.0, 0.0, 0.5],                  # on the thick stem's axis
        [2.1, 0.0, 0.3],                  # inside thick stem, near the twig
        [0.2, 0.0, 1.0 + 5e-7],           # |d| just inside A's 1e-6 tolerance
        [0.2, 0.0, 1.0 + 5e-4],           # between A's and B's tolerance
        [0.05, 0.0, 1.0 + 5e-4],          # same, interior radius (A and B differ here)
        [0.2, 0.0, 1.0 + 5e-3],           # beyond both
        [5.1, 5.2, 5.1],
        [-3.0, 4.0, 9.0],                 # far from everything
        [0.0, 0.0, 0.0],                  # exactly a cylinder start point
        [0.25, 0.0, 1.25],                # near branch mid
    ]
    rng = np.random.default_rng(41)
    extra = np.concatenate([rng.uniform(-0.6, 0.9, (150, 3)) + [0, 0, 0.5],
                            rng.uniform(-0.5, 0.5, (150, 3)) + [2, 0, 0.5]])
    cloud = np.concatenate([np.asarray(pts, np.float64), extra]).astype(np.float32)
    run_case("adversarial", q, cloud)

    # 5. a zero-length cylinder in the table (A: NaN for every point; B: sphere shell)
    rows2 = rows[:3] + [[1, 1, 1, 1, 1, 1, 0.05]] + rows[3:]
    q = qsm_from_rows(rows2, ids=[3, 1, 4, 15, 9, 2, 6])
    run_case("zero_length", q, cloud[:60])

    # 6. a single cylinder and a single point (the M=1 regime of the QSM-fitting call site)
    q = qsm_from_rows([[0.1, -0.2, 0.3, 0.4, 0.5, 1.1, 0.07]], ids=[42])
    cloud = (rng.normal(0, 0.4, (64, 3)) + [0.2, 0.1, 0.7]).astype(np.float32)
    run_case("single_cylinder", q, cloud)


if __name__ == "__main__":
    main()
