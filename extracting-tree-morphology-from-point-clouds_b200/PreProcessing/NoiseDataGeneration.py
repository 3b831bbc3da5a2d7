"""Drop-in for the reference's ``PreProcessing/NoiseDataGeneration.py``: noisy surface clouds on QSMs, the input of the
labelling path (``noiseGeneration`` reference :14-106, CLI :111-130).

Same callable, same files written (``<a>_<b>.npy`` from ``<a>_<b>_*.csv``, (N,3) float64).  The per-cylinder part — point
counts from the height-dependent density, Rodrigues rotations — is O(M) and stays in numpy float64 (``cylinder_plan``), so
the number of points per cylinder is the reference's to the last point; the O(N) part (variates, local coordinates,
rotation, translation) runs on the device (``tm_noise_cloud``).

One deliberate difference: the reference draws from numpy's global Mersenne-Twister, a sequential stream; the device draws
from a counter-based generator (Philox4x32-10 keyed by a 64-bit seed, one counter per point).  The clouds follow the same
distributions (angle uniform, axial position uniform, radial noise lognormal(-3, 0.85)) but not the same bits.  The seed is
taken from numpy's global generator, so ``np.random.seed(k)`` before the call makes a run repeatable, as it does for the
reference.  ``noise_cloud(..., variates=(theta, z, noise))`` replays given draws instead (the parity tests feed it the
reference's own).
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass

import numpy as np
import pandas as pd
import torch

from .. import api
from ..Modules.Utils import get_device

POINTS_PER_M2 = 50          # reference :40


@dataclass
class CylinderPlan:
    records: np.ndarray       # (M,14) float64: start, rotation (row-major), radius, length
    counts: np.ndarray        # (M,) int64
    first_point: np.ndarray   # (M+1,) int64 exclusive prefix of counts

    @property
    def n_points(self) -> int:
        return int(self.first_point[-1])


def _rotations_from_z(unit: np.ndarray) -> np.ndarray:
    """Rodrigues matrices taking +z to each unit axis, with the reference's treatment of the aligned case (:77-96)."""
    z_hat = np.array([0, 0, 1])
    cross = np.cross(z_hat, unit)
    sin_a = np.linalg.norm(cross, axis=1)
    cos_a = np.dot(z_hat, unit.T)
    cross[sin_a.flatten() == 0] = np.array([1, 0, 0])
    skew = np.zeros((len(unit), 3, 3))
    skew[:, 0, 1], skew[:, 0, 2] = -cross[:, 2], cross[:, 1]
    skew[:, 1, 0], skew[:, 1, 2] = cross[:, 2], -cross[:, 0]
    skew[:, 2, 0], skew[:, 2, 1] = -cross[:, 1], cross[:, 0]
    gain = ((1 - cos_a) / (sin_a ** 2 + 1e-8))[:, None, None]
    return np.eye(3)[None, :, :] + skew + np.einsum("nij,njk->nik", skew, skew) * gain


def cylinder_plan(cylinders: pd.DataFrame) -> CylinderPlan:
    """Per-cylinder quantities of the reference (:33-58, :77-96), float64 numpy in its order of operations."""
    start = cylinders[["startX", "startY", "startZ"]].values
    end = cylinders[["endX", "endY", "endZ"]].values
    radius = cylinders["radius"].values
    axis = end - start
    length = np.linalg.norm(axis, axis=1)
    with np.errstate(all="ignore"):
        unit = axis / length[:, None]
        floor_z = np.min(np.minimum(start[:, 2], end[:, 2]))
        top_z = np.max(np.maximum(start[:, 2], end[:, 2]))
        rel_height = (np.mean([start[:, 2], end[:, 2]], axis=0) - floor_z) / (top_z - floor_z)
        density = POINTS_PER_M2 * (1 - (3 / 4) * rel_height ** 0.33)          # 1 at the ground, 1/4 at the top
        counts = (2 * np.pi * radius * density).astype(int) * (length * density).astype(int)
        rot = _rotations_from_z(unit)
    if (counts < 0).any():
        raise ValueError("repeats may not contain negative values.")           # what np.repeat (:60) raises
    m = len(counts)
    records = np.empty((m, 14), dtype=np.float64)
    records[:, 0:3] = start
    records[:, 3:12] = rot.reshape(m, 9)
    records[:, 12] = radius
    records[:, 13] = length
    first = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(counts, out=first[1:])
    return CylinderPlan(records, counts.astype(np.int64), first)


def noise_cloud(cylinders: pd.DataFrame | CylinderPlan, device=None, seed: int | None = None, variates=None, as_numpy: bool = True,
                want_f32: bool = False):
    """The noisy cloud of one QSM → (N,3) float64 (numpy, or device tensors with ``as_numpy=False``)."""
    plan = cylinders if isinstance(cylinders, CylinderPlan) else cylinder_plan(cylinders)
    eng = api.get_engine(device)
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    n = plan.n_points
    if n == 0:
        empty = np.empty((0, 3)) if as_numpy else torch.empty((0, 3), dtype=torch.float64, device=eng.device)
        return (empty, empty.astype(np.float32) if as_numpy else empty.float()) if want_f32 else empty
    rec = torch.from_numpy(plan.records).to(eng.device)
    first = torch.from_numpy(plan.first_point).to(eng.device)
    out = eng.noise_cloud(rec, first, n=n, seed=seed, variates=variates, want_f32=want_f32)
    if not as_numpy:
        return out
    if want_f32:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out.cpu().numpy()


def noise_cloud_sharded(cylinders: pd.DataFrame | CylinderPlan | None, device=None, seed: int = 0, src: int = 0):
    """One process per GPU (torchrun): rank ``src`` holds the QSM, every rank generates its own contiguous rows of the cloud
    (``sharding.noise_cloud_sharded``).  Returns (device tensor (k,3) float64, (lo, hi))."""
    from .. import sharding
    eng = api.get_engine(device)
    plan = None
    if cylinders is not None:
        plan = cylinders if isinstance(cylinders, CylinderPlan) else cylinder_plan(cylinders)

    def rows(rec, first, lo, hi, sd):
        return eng.noise_cloud(rec.contiguous(), first.contiguous(), n=hi - lo, point0=lo, seed=sd)

    return sharding.noise_cloud_sharded(rows, plan.records if plan else None, plan.first_point if plan else None, seed, eng.device, src)


def noiseGeneration(data_root, npy_root):
    """
    inputs:
        data_root (str): Path to the folder where the QSMs are stored
        npy_path (str): Path to the folder where the noisy clouds should go
    """
    device = get_device()
    for name in os.listdir(data_root):
        if not name.endswith(".csv"):
            continue
        target = os.path.join(npy_root, "_".join(name.split("_")[:2]) + ".npy")      # "33_22_000000.csv" -> "33_22.npy"
        cylinders = pd.read_csv(os.path.join(data_root, name))
        cylinders.columns = cylinders.columns.str.strip()
        np.save(target, noise_cloud(cylinders, device))


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Create noise point clouds from cylinder models.")
    parser.add_argument("--cylinderDir", type=str, default=os.path.join("data", "raw", "QSM", "detailed"),
                        help="Directory containing the QSM cylinder CSV files.")
    parser.add_argument("--labelDir", type=str, default=os.path.join("data", "noised", "cloud"),
                        help="Directory where the noisy clouds should be stored.")
    args = parser.parse_args()
    noiseGeneration(data_root=os.path.join(os.getcwd(), args.cylinderDir), npy_root=os.path.join(os.getcwd(), args.labelDir))
