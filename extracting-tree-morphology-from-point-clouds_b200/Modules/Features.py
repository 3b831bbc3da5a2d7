"""The two features the hot-path drivers append after labelling (``Modules/Features.py:178-229`` with
``use_densities = use_curvatures = use_distances = use_verticalities = False``): k=15 normals and the
relative height.  Host-side and vectorised (the reference loops over points in Python); SURVEY.md §8(f)
lists a GPU version as the next step after the nearest-cylinder path.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def compute_height(points):
    """(z - z_min) / (z_max - z_min)   (Features.py:31-40)."""
    z = points[:, 2]
    lo, hi = np.min(z), np.max(z)
    return (z - lo) / (hi - lo)


def compute_normals_ckdtree(points, k=10, chunk=200_000):
    """Per point: covariance of the k nearest neighbours (relative to the point, ``np.cov`` normalisation),
    SVD, and ``v[:, -1]`` of numpy's third return value — the reference indexes the *transposed* factor
    that way (Features.py:126-131), which is kept because downstream models were trained on it."""
    n = points.shape[0]
    tree = cKDTree(points)
    out = np.zeros((n, 3))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        _, nn = tree.query(points[lo:hi], k=k)
        nb = points[nn] - points[lo:hi, None, :]                    # (c, k, 3)
        nb = nb - nb.mean(axis=1, keepdims=True)
        cov = np.einsum("nki,nkj->nij", nb, nb) / (k - 1)
        _, _, vh = np.linalg.svd(cov)
        out[lo:hi] = vh[:, :, -1]
    return out


def add_features(labeled_cloud, use_normals=True, use_heights=True, use_densities=True, use_verticalities=True,
                 use_distances=True, use_curvatures=True):
    """Append feature columns to an ``(N, 7)`` labelled cloud.  Only the features the label / projection
    drivers request (normals, heights) are implemented here; asking for the others raises."""
    if use_densities or use_verticalities or use_distances or use_curvatures:
        raise NotImplementedError("only use_normals / use_heights are part of the nearest-cylinder path's drivers")
    pts = labeled_cloud[:, :3]
    cols = [labeled_cloud]
    if use_normals:
        cols.append(compute_normals_ckdtree(pts, k=15))
    if use_heights:
        cols.append(compute_height(pts)[:, None])
    return np.concatenate(cols, axis=1)
