"""Feature step that follows the labelling in both drivers (reference Modules/Features.py:178-229): GPU neighbour search +
covariance, LAPACK decompositions on the host, against what the UNMODIFIED reference produced for the same cloud
(tests/golden/features.npz, made by tests/golden/make_golden_features.py).

Tolerances: the neighbour sets must be identical; the covariance differs from np.cov's BLAS product in the last bits, so
normals / verticality agree to 1e-6 and curvature to 1e-9 relative; density, height and distance are exact.
"""
from __future__ import annotations

import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "features.npz")


def _load():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def test_golden_layout_and_host_side_columns():
    """Column order of add_features (:178-229) and the features that are plain numpy (no GPU needed)."""
    import importlib
    g = _load()
    F = importlib.import_module("treemorph_b200.Modules.Features")
    cloud, full = g["cloud"], g["all_features"]
    assert full.shape == (len(cloud), 15) and np.array_equal(full[:, :7], cloud)
    pts = cloud[:, :3]
    assert np.array_equal(F.compute_height(pts), full[:, 12])
    assert np.array_equal(F.compute_verticality(full[:, 7:10]), full[:, 13])
    assert np.array_equal(F.compute_distance_to_center(pts), full[:, 14])
    assert np.array_equal(g["driver_features"][:, 7:10], full[:, 7:10]) and np.array_equal(g["driver_features"][:, 10], full[:, 12])


@pytest.mark.gpu
def test_neighbour_sets_and_covariance():
    import torch
    from treemorph_b200 import api
    g = _load()
    pts = g["cloud"][:, :3]
    eng = api.get_engine(torch.device("cuda", 0))
    cov, idx = eng.knn_covariance(pts, 15, want_idx=True)
    cov, idx = cov.cpu().numpy(), idx.cpu().numpy()
    assert np.array_equal(np.sort(idx, axis=1), np.sort(g["nn15"], axis=1)), "k nearest neighbours differ from cKDTree's"
    assert np.array_equal(idx[:, 0], np.arange(len(pts)))            # the point itself comes first (distance 0)
    nb = pts[g["nn15"]] - pts[:, None, :]
    want = np.stack([np.cov(x.T) for x in nb])
    assert np.allclose(cov, want, rtol=1e-11, atol=1e-18)
    # radius count against a brute-force count
    sub = np.random.default_rng(1).choice(len(pts), 400, replace=False)
    cnt = eng.radius_count(pts, 0.1).cpu().numpy()
    d2 = ((pts[sub, None, :] - pts[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(cnt[sub], (d2 <= 0.1 * 0.1).sum(1))
    with pytest.raises(ValueError):
        eng.knn_covariance(pts[:10], 15)
    bad = pts.copy()
    bad[5, 1] = np.nan
    with pytest.raises(ValueError):
        eng.knn_covariance(bad, 15)


@pytest.mark.gpu
def test_add_features_matches_reference():
    from treemorph_b200.Modules import Features as F
    g = _load()
    cloud, ref = g["cloud"], g["all_features"]
    got = F.add_features(cloud)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got[:, :7], ref[:, :7])
    close = np.isclose(got[:, 7:10], ref[:, 7:10], atol=1e-6, rtol=0).all(axis=1)
    assert close.mean() >= 0.999, f"normals differ on {np.count_nonzero(~close)} rows"
    assert np.allclose(got[:, 10], ref[:, 10], rtol=1e-9, atol=1e-15), "curvature"
    assert np.array_equal(got[:, 11], ref[:, 11]), "density"
    assert np.array_equal(got[:, 12], ref[:, 12]), "height"
    assert np.isclose(got[:, 13], ref[:, 13], atol=1e-6, rtol=0).mean() >= 0.999, "verticality"
    assert np.array_equal(got[:, 14], ref[:, 14]), "distance to centre"
    drv = F.add_features(cloud, use_densities=False, use_curvatures=False, use_distances=False, use_verticalities=False)
    assert drv.shape == g["driver_features"].shape and np.array_equal(drv[:, 10], g["driver_features"][:, 10])
    assert np.isclose(drv[:, 7:10], g["driver_features"][:, 7:10], atol=1e-6, rtol=0).all(axis=1).mean() >= 0.999
    vo = F.add_features(cloud, use_normals=False, use_heights=False, use_densities=False, use_curvatures=False, use_distances=False)
    assert vo.shape == g["verticality_only"].shape
    assert np.isclose(vo[:, 7], g["verticality_only"][:, 7], atol=1e-6, rtol=0).mean() >= 0.999
