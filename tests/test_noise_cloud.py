"""Noisy surface clouds on a QSM (reference PreProcessing/NoiseDataGeneration.py:14-106).

Pinned against tests/golden/noise.npz: clouds the UNMODIFIED reference wrote after ``np.random.seed(k)``
(tests/golden/make_golden_noise.py).  The reference's variates come from numpy's legacy global generator, whose stream is
frozen by numpy's compatibility policy, so the tests regenerate them from the seed.

Tolerances: the oracle reproduces the reference bit for bit.  The device evaluates the same float64 expressions in the same
order, but CUDA's sin/cos/log/exp are not glibc's (documented error <= 2 ulp): with coordinates of a few metres the
device cloud agrees to 1e-12 m, and that is the bound asserted.  Point counts, cylinder ownership and the counter-based
uniform variates are integers / exact and must be identical.
"""
from __future__ import annotations

import io
import os

import numpy as np
import pandas as pd
import pytest

from oracle import noise_cloud as nc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "noise.npz")
TOL_M = 1e-12


def _cases():
    with np.load(GOLDEN) as z:
        g = {k: z[k] for k in z.files}
    for name in g["names"]:
        df = pd.read_csv(io.StringIO(str(g[f"{name}__csv"])))
        df.columns = df.columns.str.strip()
        yield str(name), df, int(g[f"{name}__seed"]), g[f"{name}__cloud"], str(g[f"{name}__file"]), str(g[f"{name}__written"])


def _oracle_plan(df):
    return nc.plan(df[["startX", "startY", "startZ"]].values, df[["endX", "endY", "endZ"]].values, df["radius"].values)


def test_philox_known_answers():
    """Random123's known-answer vectors for philox4x32-10 (kat_vectors: zeros, ones, digits of pi)."""
    ctr = np.array([[0, 0, 0, 0], [0xFFFFFFFF] * 4, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], dtype=np.uint32)
    keys = [(0, 0), (0xFFFFFFFF, 0xFFFFFFFF), (0xA4093822, 0x299F31D0)]
    want = [[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8], [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD],
            [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]]
    for c, k, w in zip(ctr, keys, want):
        assert [int(x) for x in nc.philox4x32_10(c[None, :], k)[0]] == w


def test_oracle_reproduces_the_reference_bit_for_bit():
    for name, df, seed, cloud, _, _ in _cases():
        plan = _oracle_plan(df)
        np.random.seed(seed)
        theta, z, noise = nc.legacy_variates(plan)
        assert np.array_equal(nc.place(plan, theta, z, noise), cloud), name


def test_host_plan_equals_oracle_plan_and_file_names():
    from treemorph_b200.PreProcessing import NoiseDataGeneration as N
    for name, df, _, cloud, fname, written in _cases():
        want, got = _oracle_plan(df), N.cylinder_plan(df)
        assert np.array_equal(got.counts, want.count) and got.n_points == len(cloud), name
        assert np.array_equal(got.records[:, 3:12], want.rot.reshape(-1, 9))
        assert np.array_equal(got.records[:, :3], want.start) and np.array_equal(got.records[:, 13], want.length)
        assert np.array_equal(got.first_point[1:], np.cumsum(want.count)) and got.first_point[0] == 0
        assert "_".join(fname.split("_")[:2]) + ".npy" == written
    bad = next(_cases())[1].copy()
    bad.loc[3, "radius"] = -0.2                                   # negative ring count x positive height count
    with pytest.raises(ValueError):
        N.cylinder_plan(bad)


def test_counter_variates_have_the_reference_distributions():
    """uniform(0, 2 pi), uniform(0, L), lognormal(-3, 0.85) (:64-68): moments of the Philox-driven variates."""
    _, df, _, _, _, _ = next(_cases())
    plan = _oracle_plan(df)
    plan.count[:] = 5000
    theta, z, noise = nc.philox_variates(plan, seed=77)
    n = len(theta)
    assert abs(theta.mean() - np.pi) < 4 * (2 * np.pi / np.sqrt(12 * n)) and 0 <= theta.min() and theta.max() < 2 * np.pi
    u = z / plan.length[nc.owners(plan)]
    assert abs(u.mean() - 0.5) < 4 / np.sqrt(12 * n) and u.min() >= 0 and u.max() < 1
    g = (np.log(noise) + 3.0) / 0.85
    assert abs(g.mean()) < 4 / np.sqrt(n) and abs(g.std() - 1) < 0.01 and abs(np.median(noise) - np.exp(-3)) < 1e-3
    assert abs(np.corrcoef(theta, noise)[0, 1]) < 0.01 and abs(np.corrcoef(u, g)[0, 1]) < 0.01


# ---- device -----------------------------------------------------------------------------------------
def _engine():
    import torch
    from treemorph_b200 import api
    return api.get_engine(torch.device("cuda", 0))


@pytest.mark.gpu
def test_replaying_the_reference_draws_reproduces_its_cloud():
    from treemorph_b200.PreProcessing import NoiseDataGeneration as N
    for name, df, seed, cloud, _, _ in _cases():
        plan = _oracle_plan(df)
        np.random.seed(seed)
        variates = nc.legacy_variates(plan)
        got, got32 = N.noise_cloud(df, "cuda:0", variates=variates, want_f32=True)
        assert got.shape == cloud.shape and got.dtype == np.float64
        assert np.abs(got - cloud).max() <= TOL_M, (name, np.abs(got - cloud).max())
        assert np.array_equal(got32, got.astype(np.float32))


@pytest.mark.gpu
def test_counter_based_cloud_equals_the_oracle_and_does_not_depend_on_sharding():
    import torch
    from treemorph_b200.PreProcessing import NoiseDataGeneration as N
    eng = _engine()
    for name, df, _, _, _, _ in _cases():
        plan = _oracle_plan(df)
        seed = 0x1234_5678_9ABC_DEF0 + len(df)
        want = nc.place(plan, *nc.philox_variates(plan, seed))
        got = N.noise_cloud(df, "cuda:0", seed=seed)
        assert np.abs(got - want).max() <= TOL_M, (name, np.abs(got - want).max())
        # rows [a, b) of the cloud from a second call: the same bits as in the whole cloud
        hp = N.cylinder_plan(df)
        rec, first = torch.from_numpy(hp.records).cuda(), torch.from_numpy(hp.first_point).cuda()
        a, b = hp.n_points // 3 + 7, hp.n_points - 5
        part = eng.noise_cloud(rec, first, n=b - a, point0=a, seed=seed).cpu().numpy()
        assert np.array_equal(part, got[a:b])
        assert not np.array_equal(N.noise_cloud(df, "cuda:0", seed=seed + 1), got)
        rows, (lo, hi) = N.noise_cloud_sharded(df, "cuda:0", seed=seed)          # one rank: the whole cloud
        assert (lo, hi) == (0, hp.n_points) and np.array_equal(rows.cpu().numpy(), got)


@pytest.mark.gpu
def test_large_cloud_properties():
    """Size-independent checks on a 2M-point cloud: every point belongs to its cylinder's slab, and its distance from the
    axis minus the radius is the lognormal noise (median exp(-3) = 5 cm, the reference's class-balancing threshold)."""
    from treemorph_b200 import synth
    from treemorph_b200.PreProcessing import NoiseDataGeneration as N
    df = synth.qsm_dataframe(synth.random_qsm(3000, seed=5))
    hp = N.cylinder_plan(df)
    scale = 2_000_000 / max(1, hp.n_points)
    hp.counts[:] = np.maximum(1, (hp.counts * scale).astype(np.int64))
    hp.first_point[1:] = np.cumsum(hp.counts)
    cloud = N.noise_cloud(hp, "cuda:0", seed=9)
    assert cloud.shape == (hp.n_points, 3) and np.isfinite(cloud).all()
    cid = np.repeat(np.arange(len(hp.counts)), hp.counts)
    start, end = df[["startX", "startY", "startZ"]].values, df[["endX", "endY", "endZ"]].values
    axis = end - start
    length = np.linalg.norm(axis, axis=1)
    unit = axis / length[:, None]
    v = cloud - start[cid]
    t = (v * unit[cid]).sum(1)
    assert (t > -1e-9).all() and (t < length[cid] + 1e-9).all()
    radial = np.linalg.norm(v - t[:, None] * unit[cid], axis=1) - df["radius"].values[cid]
    assert (radial > 0).all() and abs(np.median(radial) - np.exp(-3.0)) < 5e-4
    g = (np.log(radial) + 3.0) / 0.85
    assert abs(g.mean()) < 5e-3 and abs(g.std() - 1.0) < 5e-3


@pytest.mark.gpu
def test_driver_writes_the_reference_files(tmp_path):
    from treemorph_b200.PreProcessing import NoiseDataGeneration as N
    src, dst = tmp_path / "qsm", tmp_path / "cloud"
    src.mkdir()
    dst.mkdir()
    sizes = {}
    with np.load(GOLDEN) as z:
        for name in z["names"]:
            (src / str(z[f"{name}__file"])).write_text(str(z[f"{name}__csv"]))
            sizes[str(z[f"{name}__written"])] = z[f"{name}__cloud"].shape
    (src / "notes.txt").write_text("not a table")
    np.random.seed(5)
    N.noiseGeneration(str(src), str(dst))
    first = {f: np.load(dst / f) for f in sorted(os.listdir(dst))}
    assert {f: a.shape for f, a in first.items()} == sizes and all(a.dtype == np.float64 for a in first.values())
    np.random.seed(5)
    N.noiseGeneration(str(src), str(dst))
    assert all(np.array_equal(np.load(dst / f), a) for f, a in first.items())       # repeatable under np.random.seed


@pytest.mark.gpu
def test_edge_cases():
    import torch
    from treemorph_b200.PreProcessing import NoiseDataGeneration as N
    eng = _engine()
    _, df, _, _, _, _ = next(_cases())
    none = df.copy()
    none["radius"] = 1e-4                                        # no cylinder earns a point
    assert N.noise_cloud(none, "cuda:0").shape == (0, 3)
    rec = torch.zeros((0, 14), dtype=torch.float64, device="cuda")
    first = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert eng.noise_cloud(rec, first).shape == (0, 3)
    with pytest.raises(IndexError):                              # rows asked of an empty table (TM_ERR_NO_CYLINDERS)
        eng.noise_cloud(rec, first, n=4)
    hp = N.cylinder_plan(df)
    rec_d, first_d = torch.from_numpy(hp.records).cuda(), torch.from_numpy(hp.first_point).cuda()
    over = eng.noise_cloud(rec_d, first_d, n=300, point0=hp.n_points - 100, seed=1).cpu().numpy()     # 200 rows past the end
    assert np.isfinite(over[:100]).all() and np.isnan(over[100:]).all()
    assert np.array_equal(over[:100], eng.noise_cloud(rec_d, first_d, seed=1).cpu().numpy()[-100:])
    with pytest.raises(ValueError):
        eng.noise_cloud(torch.from_numpy(hp.records).cuda(), torch.from_numpy(hp.first_point).cuda(), n=10,
                        variates=(np.zeros(10), np.zeros(10), np.zeros(9)))
