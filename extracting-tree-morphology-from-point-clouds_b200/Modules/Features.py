"""Drop-in for the reference's ``Modules/Features.py``: the per-point features the label / projection drivers append right
after the nearest-cylinder path (``add_features``, reference :178-229, called at ``LabelGenerationCuda.py:197-198`` and
``Projection.py:420-427``).

The reference builds a scipy cKDTree and then loops over the points in Python (``np.cov`` + LAPACK per point).  Here

* the neighbour search (k nearest / radius) and the covariance run on the GPU in float64
  (``tm_knn_covariance`` / ``tm_radius_count``, csrc/tm_knn.cu), one launch per cloud;
* the 3x3 decompositions stay with LAPACK on the host, BATCHED and spread over threads: the reference's normal is
  ``v[:, -1]`` of ``np.linalg.svd``'s third return value, i.e. the last COLUMN of ``Vh`` with LAPACK's own sign choices —
  no independent eigen-solver reproduces that, the same LAPACK call on the same matrix does;
* heights, verticality and distance to the centre are the reference's own numpy expressions.

There is no CPU fallback for the neighbour search: without a CUDA device these functions raise.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import api

_CHUNK = 1 << 16


def _threads() -> int:
    try:
        return max(1, min(16, len(os.sched_getaffinity(0))))
    except AttributeError:
        return max(1, min(16, os.cpu_count() or 1))


def _batched(fn, mats: np.ndarray):
    """``fn`` (a batched numpy.linalg routine; it releases the GIL) over chunks of ``mats`` on a few threads."""
    n = mats.shape[0]
    parts = [(lo, min(n, lo + _CHUNK)) for lo in range(0, n, _CHUNK)]
    if len(parts) <= 1:
        return [fn(mats)]
    with ThreadPoolExecutor(max_workers=_threads()) as pool:
        return list(pool.map(lambda p: fn(mats[p[0]:p[1]]), parts))


def _covariances(points, k: int) -> np.ndarray:
    points = np.asarray(points)
    if points.shape[0] < k:
        raise ValueError(f"{points.shape[0]} points cannot have {k} nearest neighbours each")
    eng = api.get_engine()
    return eng.knn_covariance(points, k).cpu().numpy()


def compute_normals_ckdtree(points, k=10):
    """Reference :111-133 — ``v[:, -1]`` of ``_, _, v = svd(cov(k nearest neighbours - point))``."""
    cov = _covariances(points, k)
    return np.concatenate([vh[:, :, -1] for _, _, vh in _batched(np.linalg.svd, cov)], axis=0)


def compute_normals(points, k=10):
    """Reference :11-30 (the sklearn variant): ``v[-1]``, the last ROW of ``Vh``."""
    cov = _covariances(points, k)
    return np.concatenate([vh[:, -1, :] for _, _, vh in _batched(np.linalg.svd, cov)], axis=0)


def compute_curvature_ckdtree(points, k=10):
    """Reference :136-157 — smallest eigenvalue over (sum of eigenvalues + 1e-6)."""
    cov = _covariances(points, k)
    ev = np.concatenate(_batched(np.linalg.eigvalsh, cov), axis=0)
    return ev[:, 0] / (np.sum(ev, axis=1) + 1e-6)


compute_curvature = compute_curvature_ckdtree          # reference :77-107: same quantity through sklearn + eigh


def compute_density_ckdtree(points, radius=0.1):
    """Reference :160-172 — ``len(tree.query_ball_point(point, r=radius))`` per point."""
    eng = api.get_engine()
    return eng.radius_count(np.asarray(points), radius).cpu().numpy().astype(np.int64)


compute_density = compute_density_ckdtree              # reference :42-53


def compute_height(points):
    """Reference :31-40 — (z - z_min) / (z_max - z_min)."""
    z_min = np.min(points[:, 2])
    z_max = np.max(points[:, 2])
    return (points[:, 2] - z_min) / (z_max - z_min)


def compute_verticality(normals):
    """Reference :55-64 — |normal . z|."""
    return np.abs(np.dot(normals, np.array([0, 0, 1])))


def compute_distance_to_center(points):
    """Reference :66-75 — horizontal distance to the cloud's mean xy."""
    center_xy = np.mean(points[:, :2], axis=0)
    return np.linalg.norm(points[:, :2] - center_xy, axis=1)


def add_features(labeled_cloud, use_normals=True, use_heights=True, use_densities=True, use_verticalities=True,
                 use_distances=True, use_curvatures=True):
    """Append the selected feature columns to an ``(N, >=3)`` labelled cloud, in the reference's order (:178-229):
    normals (3), curvature, density, relative height, verticality, distance to the centre."""
    points = labeled_cloud[:, :3]
    cols = [labeled_cloud]
    normals = None
    if use_normals:
        normals = compute_normals_ckdtree(points, k=15)
        cols.append(normals)
    if use_curvatures:
        cols.append(compute_curvature_ckdtree(points, k=10)[:, np.newaxis])
    if use_densities:
        cols.append(compute_density_ckdtree(points)[:, np.newaxis])
    if use_heights:
        cols.append(compute_height(points)[:, np.newaxis])
    if use_verticalities:
        if normals is None:                       # the reference falls back to its sklearn variant here (:217)
            normals = compute_normals(points, k=15)
        cols.append(compute_verticality(normals)[:, np.newaxis])
    if use_distances:
        cols.append(compute_distance_to_center(points)[:, np.newaxis])
    return np.concatenate(cols, axis=1)
