"""Parity tests proper (need a B200): the CUDA path, called through the C-ABI library, against

* the committed golden vectors minted from the reference itself (bit-for-bit),
* the CPU oracle on seeded inputs at sizes it finishes in seconds,
* size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): cylinder indices exact except near-ties (top-2 distances within
1e-6 relative); offsets and distances within 1e-5 m.  In practice the kernels are bit-identical.
"""
import gc
import os

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import assert_same_bits, golden_names, load_golden
from helpers import assert_parity, make_case, oracle_label
from oracle import oracle as _oracle

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from treemorph_b200 import api
    from treemorph_b200.Modules import Projection as P
    from treemorph_b200.PreProcessing import LabelGenerationCuda as L


@pytest.fixture(scope="module")
def eng():
    e = api.Engine()
    yield e
    e.close()


def _install(eng, case):
    dev = eng.device
    eng.set_cylinders(torch.tensor(case["start"], device=dev), torch.tensor(case["radius"], device=dev),
                      torch.tensor(case["length"], device=dev), torch.tensor(case["unit"], device=dev),
                      torch.tensor(case["ids"], device=dev))


def _label(eng, case, pts, mode, **kw):
    res = eng.label(torch.tensor(pts, device=eng.device), api.VARIANTS[case["variant"].name], mode=mode,
                    want=("index", "id", "dist", "offset", "radius"), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in res.items()}


# ---- golden vectors ---------------------------------------------------------------------------------

@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("vn", ["A", "B"])
def test_dropin_cloud_matches_reference_golden(name, vn):
    """generate_offset_cloud_cuda_batched of both drop-in modules == the reference's (N,7) output."""
    g = load_golden(name)
    mod = L if vn == "A" else P
    out = mod.generate_offset_cloud_cuda_batched(g["cloud"], pd.DataFrame(g["qsm"]), torch.device("cuda"))
    assert out.dtype == np.float64
    assert_same_bits(out, g[f"ref_{vn}_cloud"], f"{name}/{vn}")


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("vn", ["A", "B"])
def test_dropin_kernel_matches_reference_golden(name, vn):
    """closest_cylinder_cuda_batch with tensors built the way the reference's driver builds them."""
    g = load_golden(name)
    df = pd.DataFrame(g["qsm"])
    dev = torch.device("cuda")
    start = torch.tensor(df[["startX", "startY", "startZ"]].values, dtype=torch.float32, device=dev)
    radius = torch.tensor(df["radius"].values, dtype=torch.float32, device=dev)
    ids = torch.tensor(df["ID"].values, dtype=torch.int32, device=dev)
    length = torch.tensor(g[f"ref_{vn}_length"], device=dev)
    unit = torch.tensor(g[f"ref_{vn}_unit"], device=dev)
    mod = L if vn == "A" else P
    rid, rdist, roff = mod.closest_cylinder_cuda_batch(g["cloud"][:, :3], start, radius, length, unit, ids, dev)
    assert rid.dtype == np.int32 and rdist.dtype == np.float32 and roff.dtype == np.float32
    if len(df) > 1 and start.stride(1) == 1:
        pytest.skip("torch.tensor() made the DataFrame values C-contiguous here; layout-specific rounding differs")
    assert_same_bits(rid, g[f"ref_{vn}_id"], "ids")
    assert_same_bits(rdist, g[f"ref_{vn}_dist"], "distances")
    assert_same_bits(roff, g[f"ref_{vn}_off"], "offsets")


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("mode", ["brute", "grid"])
def test_engine_modes_match_golden(eng, name, mode):
    g = load_golden(name)
    q = g["qsm"]
    m = len(q["ID"])
    for vn in "AB":
        var = api.VARIANTS[vn]
        dev = eng.device
        start = torch.tensor(np.stack([q["startX"], q["startY"], q["startZ"]], 1).astype(np.float32), device=dev)
        eng.set_cylinders(start, torch.tensor(q["radius"].astype(np.float32), device=dev),
                          torch.tensor(g[f"ref_{vn}_length"], device=dev), torch.tensor(g[f"ref_{vn}_unit"], device=dev),
                          torch.tensor(q["ID"].astype(np.int32), device=dev))
        pts = torch.tensor(g["cloud"][:, :3].astype(np.float32), device=dev)
        res = eng.label(pts, var, mode=mode, norm_fma=(m == 1), want=("id", "dist", "offset"))
        assert_same_bits(res["id"].cpu().numpy(), g[f"ref_{vn}_id"], f"{name}/{vn}/{mode} ids")
        assert_same_bits(res["dist"].cpu().numpy(), g[f"ref_{vn}_dist"], f"{name}/{vn}/{mode} distances")
        assert_same_bits(res["offset"].cpu().numpy(), g[f"ref_{vn}_off"], f"{name}/{vn}/{mode} offsets")


# ---- oracle on seeded inputs --------------------------------------------------------------------------

@pytest.mark.parametrize("vn", ["A", "B"])
@pytest.mark.parametrize("mode", ["brute", "grid"])
def test_config1_100k_points_2k_cylinders(eng, vn, mode):
    """BASELINE.json configs[0]: 100k points vs a 2k-cylinder QSM."""
    case = make_case(2000, 100_000, seed=1, variant=vn)
    ora = oracle_label(case)
    _install(eng, case)
    got = _label(eng, case, case["points"], mode)
    assert_parity(got, ora, f"config1/{vn}/{mode}", require_bitwise=True)
    assert (got["radius"] == case["radius"][got["index"]]).all()


@pytest.mark.parametrize("vn", ["A", "B"])
def test_plot_of_several_trees_grid(eng, vn):
    """A multi-tree plot (12k cylinders) with model-residual and lognormal clouds; oracle on 60k points."""
    case = make_case(12_000, 60_000, seed=7, variant=vn, id_offset=1000)
    _install(eng, case)
    for noise in ("lognormal", "model"):
        from treemorph_b200 import synth
        pts = synth.sample_points(case["qsm"], 60_000, seed=8, noise=noise)
        ora = oracle_label(case, pts)
        got = _label(eng, case, pts, "grid")
        assert_parity(got, ora, f"plot/{vn}/{noise}", require_bitwise=True)
        st = eng.stats()
        assert st["mode_used"] == 2 and st["pairs_evaluated"] < 0.05 * 60_000 * 12_000     # the grid really prunes


def test_cell_sizes_do_not_change_results(eng):
    case = make_case(3000, 50_000, seed=21, variant="A")
    _install(eng, case)
    ref = _label(eng, case, case["points"], "brute")
    for cell in (0.1, 0.25, 0.6, 1.5):
        got = _label(eng, case, case["points"], "grid", cell_size=cell)
        for k in ("index", "id", "dist", "offset"):
            assert_same_bits(got[k], ref[k], f"cell {cell}: {k}")


def test_far_and_nonfinite_points(eng):
    """Points outside the voxel grid, NaN / Inf coordinates: same answers as the oracle (A.4)."""
    case = make_case(1500, 20_000, seed=31, variant="B")
    pts = case["points"].copy()
    rng = np.random.default_rng(5)
    pts[:300] += rng.normal(0, 30, (300, 3)).astype(np.float32)          # far outside the QSM box
    pts[300] = [np.nan, 0, 0]
    pts[301] = [0, np.inf, 0]
    pts[302] = [1e30, -1e30, 1e30]
    ora = oracle_label(case, pts)
    _install(eng, case)
    with np.errstate(all="ignore"):
        for mode in ("grid", "brute"):
            got = _label(eng, case, pts, mode)
            assert_parity(got, ora, f"outliers/{mode}", require_bitwise=True)
            if mode == "grid":
                st = eng.stats()
                assert st["points_brute"] == 2 and st["points_tree"] >= 250       # NaN / Inf rows: exhaustive; the rest: tree


def test_empty_and_error_cases(eng):
    case = make_case(50, 10, seed=41)
    _install(eng, case)
    res = eng.label(torch.zeros((0, 3), device=eng.device), want=("id", "dist", "offset"))
    assert res["id"].shape == (0,) and res["offset"].shape == (0, 3)
    dev = eng.device
    eng.set_cylinders(torch.zeros((0, 3), device=dev), torch.zeros(0, device=dev), torch.zeros((0, 1), device=dev),
                      torch.zeros((0, 3), device=dev), torch.zeros(0, dtype=torch.int32, device=dev))
    with pytest.raises(IndexError):               # the reference raises from argmin on an empty dim
        eng.label(torch.zeros((4, 3), device=dev))
    eng.label(torch.zeros((0, 3), device=dev))    # N == 0 stays a no-op
    with pytest.raises(ValueError):
        eng.label(torch.zeros((4, 2), device=dev))
    out = L.generate_offset_cloud_cuda_batched(np.zeros((0, 3)), pd.DataFrame(case["qsm"]), torch.device("cuda"))
    assert out.shape == (0, 7)


def test_ragged_strided_and_f64_clouds(eng):
    """Clouds with extra columns (row stride > 3), float64 input, odd sizes; host path == device path."""
    case = make_case(800, 4097, seed=51, variant="A")
    ora = oracle_label(case)
    df = pd.DataFrame(case["qsm"])
    wide = np.concatenate([case["points"].astype(np.float64), np.random.default_rng(1).random((4097, 4))], axis=1)
    out = L.generate_offset_cloud_cuda_batched(wide, df, torch.device("cuda"), batch_size=300)
    assert_same_bits(out[:, :3], wide[:, :3], "xyz copied at the caller's precision")
    assert_same_bits(out[:, 3:6].astype(np.float32), ora["offset"], "offsets")
    assert_same_bits(out[:, 6].astype(np.int32), ora["id"], "ids")
    _install(eng, case)
    wide32 = torch.tensor(wide.astype(np.float32), device=eng.device)
    got = eng.label(wide32, api.VARIANT_A, want=("id", "offset"))
    assert_same_bits(got["offset"].cpu().numpy(), ora["offset"], "strided device points")


def test_third_call_site_regime_small_m(eng):
    """QSMFittingDepthFirst.py:1079-1081: a handful of cylinders as C-ordered tensors, move_points_to_mantle=True,
    only `distances < eps` is consumed."""
    case = make_case(6, 5000, seed=61, variant="B")
    dev = torch.device("cuda")
    args = [torch.tensor(case[k], device=dev) for k in ("start", "radius", "length", "unit", "ids")]
    rid, rdist, roff = P.closest_cylinder_cuda_batch(case["points"], *args, dev, move_points_to_mantle=True)
    ora = oracle_label(case, norm_fma=True)            # contiguous tensors → ATen's fused norm
    assert_same_bits(rid, ora["id"], "ids")
    assert_same_bits(rdist, ora["dist"], "distances")
    assert_same_bits(roff, ora["offset"], "offsets")
    assert ((rdist < 0.1) == (ora["dist"] < 0.1)).all()


def test_arithmetic_selftest(eng):
    """div3 / sqrt_rn of the pair evaluation are bit-identical to the compiler's IEEE __fdiv_rn / __fsqrt_rn."""
    assert eng.selftest_arithmetic(1 << 25, seed=7) == 0
    assert eng.selftest_arithmetic(1 << 22, seed=12345) == 0


def _append_cylinders(case, rows):
    """rows: (start xyz, unit xyz, length, radius) appended to the case's fp32 kernel inputs."""
    rows = np.asarray(rows, dtype=np.float32)
    out = dict(case)
    out["start"] = np.concatenate([case["start"], rows[:, 0:3]])
    out["unit"] = np.concatenate([case["unit"], rows[:, 3:6]])
    out["length"] = np.concatenate([case["length"], rows[:, 6:7]])
    out["radius"] = np.concatenate([case["radius"], rows[:, 7]])
    out["ids"] = np.concatenate([case["ids"], np.arange(len(rows), dtype=np.int32) + 900000])
    return out


@pytest.mark.parametrize("vn", ["A", "B"])
def test_axis_parallel_cylinders_far_away(eng, vn):
    """Variant A gives NaN (which wins the argmin) to every point exactly on the axis LINE of an axis-parallel
    cylinder, however far away; pruning must not lose those.  Variant B has no NaN; it must match as well."""
    case = _append_cylinders(make_case(3000, 30_000, seed=71, variant=vn),
                             [[7.0, -3.0, 40.0, 0, 0, 1, 1.0, 0.1],        # vertical, 40 m above the tree
                              [-25.0, 0.5, 2.0, 1, 0, 0, 0.7, 0.05],       # along x, 25 m to the side
                              [0.25, 30.0, 1.5, 0, -1, 0, 0.5, 0.02],      # along -y
                              [7.0, -3.0, 44.0, 0, 0, 1, 1.0, 0.1]])       # same line as the first: lowest index wins
    pts = case["points"].copy()
    pts[:200, 0], pts[:200, 1] = 7.0, -3.0                                  # on the line of rows M and M+3
    pts[200:400, 1], pts[200:400, 2] = 0.5, 2.0                             # on the line of row M+1
    pts[400:500, 0], pts[400:500, 2] = 0.25, 1.5                            # on the line of row M+2
    ora = oracle_label(case, pts)
    if vn == "A":
        assert np.isnan(ora["dist"][:500]).all() and (ora["index"][:200] == 3000).all()
    _install(eng, case)
    with np.errstate(all="ignore"):
        for mode in ("grid", "brute"):
            got = _label(eng, case, pts, mode)
            assert_parity(got, ora, f"axis-parallel/{vn}/{mode}", require_bitwise=True)


@pytest.mark.parametrize("vn", ["A", "B"])
def test_long_special_and_remote_cylinders(eng, vn):
    """Cylinders the voxel index cannot list the usual way: one spanning > 32k voxels (long list), one with a
    non-unit axis and one with NaN entries (special list: evaluated for every point), plus points that sit in
    voxels with empty tiles and far outside the grid (both: tree search)."""
    base = make_case(2500, 40_000, seed=81, variant=vn)
    d = np.array([1.0, 1.0, 1.0], np.float32) / np.sqrt(np.float32(3))
    rows = [[-6.0, -6.0, 0.0, d[0], d[1], d[2], 20.8, 0.05],                # 20.8 m diagonal: long list
            [1.0, 1.0, 3.0, 0.5, 0.0, 0.0, 1.0, 0.1],                       # |u| = 0.5: special
            [-2.0, 2.0, 2.0, 0.0, 0.6, 0.8, 0.4, 0.03]]
    case = _append_cylinders(base, rows)
    rng = np.random.default_rng(82)
    t = rng.uniform(0, 20.8, 4000).astype(np.float32)
    along = np.array([-6.0, -6.0, 0.0], np.float32) + t[:, None] * d + rng.normal(0, 0.05, (4000, 3)).astype(np.float32)
    remote = (rng.uniform(-1, 1, (3000, 3)) * [9, 9, 3] + [0, 0, 14]).astype(np.float32)      # inside the grid, empty tiles
    far = (rng.normal(0, 60, (500, 3))).astype(np.float32)                                      # mostly outside the grid
    pts = np.concatenate([case["points"], along, remote, far])
    ora = oracle_label(case, pts)
    _install(eng, case)
    st = None
    for mode in ("grid", "brute"):
        got = _label(eng, case, pts, mode)
        st = st or eng.stats()
        assert_parity(got, ora, f"long+special/{vn}/{mode}", require_bitwise=True)
    # a NaN row in the table: variant A answers every point with it, variant B too where NaN propagates
    case2 = _append_cylinders(base, [[np.nan, 0.0, 0.0, 0, 0, 1, 1.0, 0.1]])
    ora2 = oracle_label(case2, pts[:20_000])
    _install(eng, case2)
    with np.errstate(all="ignore"):
        got2 = _label(eng, case2, pts[:20_000], "grid")
        assert_parity(got2, ora2, f"nan-row/{vn}", require_bitwise=True)
    assert st["mode_used"] == 2 and st["points_far"] > 0 and st["points_ring"] + st["points_tree"] >= 3000 and st["points_tree"] > 0
    assert st["points_brute"] == 0
    assert st["points_grid"] + st["points_far"] + st["points_ring"] + st["points_tree"] + st["points_brute"] == len(pts)


@pytest.mark.parametrize("vn", ["A", "B"])
def test_crowded_voxels_and_noise_clouds(eng, vn):
    """Tiles longer than the in-kernel sort (thousands of twigs in one voxel: unsorted fallback, many staging
    chunks), a dense cluster of duplicates, and a cloud that is mostly far from every cylinder (far part of the
    tiles, tree search with and without an incumbent)."""
    base = make_case(600, 20_000, seed=61, variant=vn)
    rng = np.random.default_rng(62)
    k = 3000
    centre = np.array([1.1, 0.6, 2.3])
    p0 = centre + rng.uniform(-0.1, 0.1, (k, 3))
    d = rng.normal(size=(k, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rows = [[*p0[i], *d[i], rng.uniform(0.005, 0.04), rng.uniform(0.001, 0.004)] for i in range(k)]
    rows += [[*centre, 0.0, 0.6, 0.8, 0.05, 0.002]] * 40                                   # 40 identical twigs: lowest row wins
    case = _append_cylinders(base, rows)
    near = (centre + rng.normal(0, 0.08, (30_000, 3))).astype(np.float32)
    haze = (rng.uniform(-1, 1, (40_000, 3)) * [6, 6, 6] + [0, 0, 6]).astype(np.float32)      # 0 .. several metres from the tree
    pts = np.concatenate([case["points"], near, haze])
    ora = oracle_label(case, pts)
    _install(eng, case)
    for mode in ("grid", "brute"):
        got = _label(eng, case, pts, mode)
        assert_parity(got, ora, f"crowded/{vn}/{mode}", require_bitwise=True)
        if mode == "grid":
            st = eng.stats()
            assert st["points_far"] > 0 and st["points_tree"] > 30_000 and st["points_ring"] == 0      # many stragglers: tree search
            assert st["points_grid"] + st["points_far"] + st["points_ring"] + st["points_tree"] + st["points_brute"] == len(pts)


@pytest.mark.parametrize("shift", [(1000.0, -2500.0, 300.0), (4.0e5, 5.6e6, 120.0)])
def test_projected_coordinates(eng, shift):
    """Plots in projected (national grid / UTM-like) coordinates: fp32 world coordinates are coarse there (1e-4 .. 0.5 m
    per ulp), the reference's answers are what they are, and the pruning must still reproduce them exactly."""
    from treemorph_b200 import synth
    q = synth.random_qsm(1200, seed=91)
    for k in ("startX", "endX"):
        q[k] = q[k] + shift[0]
    for k in ("startY", "endY"):
        q[k] = q[k] + shift[1]
    for k in ("startZ", "endZ"):
        q[k] = q[k] + shift[2]
    pts = synth.sample_points(q, 30_000, seed=92)
    start, radius, length, unit, ids = synth.cylinder_arrays(q)
    case = {"start": start, "radius": radius, "length": length, "unit": unit, "ids": ids, "variant": _oracle.VARIANT_A, "points": pts}
    with np.errstate(all="ignore"):
        ora = oracle_label(case, pts)
        _install(eng, case)
        for mode in ("grid", "brute"):
            got = _label(eng, case, pts, mode)
            assert_parity(got, ora, f"shifted{shift}/{mode}", require_bitwise=True)


@pytest.mark.parametrize("f64", [False, True])
def test_host_pipeline_variants_agree(eng, f64, monkeypatch):
    """tm_label_cloud_host: host-assembled records (16 B/point over PCIe) and device-assembled records (56 B/point) are
    the same bits, for pageable clouds with extra columns, several chunks and a ragged tail, and equal the oracle."""
    case = make_case(900, 70_001, seed=71, variant="B")
    _install(eng, case)
    pts = case["points"]
    cloud = np.concatenate([pts.astype(np.float64) if f64 else pts, np.full((len(pts), 2), 7, pts.dtype if not f64 else np.float64)], axis=1)
    ora = oracle_label(case, pts)
    monkeypatch.setenv("TM_HOST_CHUNK", "16384")
    outs = {}
    for mode in ("0", "3", "16"):
        monkeypatch.setenv("TM_HOST_ASSEMBLE", mode)
        rec, dist = eng.label_cloud_host(cloud, api.VARIANT_B, mode="grid", want_dist=True)
        info = eng.host_pipeline_info()
        assert info["host_threads"] == int(mode) and info["d2h_bytes_per_point"] == (60 if mode == "0" else 20)
        outs[mode] = (rec, dist)
        assert np.array_equal(rec[:, :3], cloud[:, :3].astype(np.float64))
        assert np.array_equal(rec[:, 3:6], ora["offset"].astype(np.float64), equal_nan=True)
        assert np.array_equal(rec[:, 6], ora["id"].astype(np.float64))
        assert np.array_equal(dist, ora["dist"], equal_nan=True)
    assert np.array_equal(outs["0"][0], outs["3"][0], equal_nan=True) and np.array_equal(outs["0"][0], outs["16"][0], equal_nan=True)
    # page-locked record array: some chunks are assembled on the device and written by the copy engine, the others by
    # the host workers (TM_HOST_SPLIT percent / the rest); every mix gives the same rows
    monkeypatch.setenv("TM_HOST_ASSEMBLE", "4")
    for split, per_point in (("0", 20), ("50", 40), ("100", 60), (None, 20)):
        if split is None:
            monkeypatch.delenv("TM_HOST_SPLIT")                              # off unless asked for
        else:
            monkeypatch.setenv("TM_HOST_SPLIT", split)
        out = torch.full((len(pts), 7), -1.0, dtype=torch.float64).pin_memory().numpy()
        rec, dist = eng.label_cloud_host(cloud, api.VARIANT_B, mode="grid", want_dist=True, out=out)
        want = 20 if f64 else per_point              # float64 clouds: the device only has their float32 rounding, hosts assemble all
        assert rec is out and abs(eng.host_pipeline_info()["d2h_bytes_per_point"] - want) <= 6
        assert np.array_equal(rec, outs["0"][0], equal_nan=True) and np.array_equal(dist, ora["dist"], equal_nan=True)
    for pinned in ("1", "0"):                                                # result array allocated by the engine
        monkeypatch.setenv("TM_PINNED_OUT", pinned)
        rec = eng.label_cloud_host(cloud, api.VARIANT_B, mode="grid")
        assert torch.from_numpy(rec).is_pinned() == (pinned == "1")
        assert np.array_equal(rec, outs["0"][0], equal_nan=True)
    monkeypatch.setenv("TM_PINNED_OUT", "1")
    live = api._pinned_live
    rec = eng.label_cloud_host(cloud, api.VARIANT_B, mode="grid")
    assert api._pinned_live == live + rec.nbytes
    del rec
    gc.collect()
    assert api._pinned_live == live                                          # freed arrays give their budget back
    monkeypatch.setenv("TM_PINNED_OUT_TOTAL_MB", "0")
    assert not torch.from_numpy(eng.label_cloud_host(cloud, api.VARIANT_B, mode="grid")).is_pinned()


def _fuzz_case(rng):
    """Random cylinder tables well outside the synthetic-tree regime: scales from millimetres to tens of metres, needles
    and discs, clusters, zero-length / zero-radius / duplicated cylinders, optional NaN rows; points near, far, on axes."""
    m = int(rng.choice([1, 2, 5, 37, 300, 2500]))
    scale = float(10.0 ** rng.uniform(-4.0, 1.3))
    centre = rng.normal(0, 1, 3) * float(10.0 ** rng.uniform(-1, 2.5))
    start = centre + rng.normal(0, 1, (m, 3)) * scale * rng.choice([0.3, 3.0, 30.0])
    d = rng.normal(size=(m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    length = np.abs(rng.normal(0, 1, m)) * scale * rng.choice([0.1, 1.0, 10.0])
    radius = np.abs(rng.normal(0, 1, m)) * scale * rng.choice([0.01, 0.2, 2.0])
    if m > 4:
        length[rng.integers(m)] = 0.0                      # zero-length cylinder (NaN unit in variant A)
        radius[rng.integers(m)] = 0.0
        k = rng.integers(1, m)
        start[k], d[k], length[k], radius[k] = start[0], d[0], length[0], radius[0]      # exact duplicate: lowest row wins
        d[rng.integers(m)] = [0.0, 0.0, 1.0]               # axis-parallel
    end = start + d * length[:, None]
    n = int(rng.choice([1, 33, 1000, 20_000]))
    pick = rng.integers(0, m, n)
    t = rng.uniform(-0.3, 1.3, n)
    base = start[pick] + (end[pick] - start[pick]) * t[:, None]
    noise = rng.normal(size=(n, 3)) * (radius[pick][:, None] + scale * rng.choice([0.01, 0.3, 5.0]))
    pts = base + noise
    on_axis = rng.random(n) < 0.05
    pts[on_axis] = base[on_axis]                           # exactly on an axis line / inside a cylinder
    far = rng.random(n) < 0.05
    pts[far] += rng.normal(size=(int(far.sum()), 3)) * scale * 300
    return start.astype(np.float32), end.astype(np.float32), radius.astype(np.float32), pts.astype(np.float32)


@pytest.mark.parametrize("seed", range(48))
def test_fuzz_grid_equals_exhaustive_equals_oracle(eng, seed):
    rng = np.random.default_rng(1000 + seed)
    start, end, radius, pts = _fuzz_case(rng)
    vn = "AB"[seed % 2]
    var = _oracle.VARIANTS[vn]
    with np.errstate(all="ignore"):
        length, unit = _oracle.prepare(start, end, var)
    ids = (rng.permutation(len(start)) + 5).astype(np.int32)
    case = {"start": start, "radius": radius, "length": length, "unit": unit, "ids": ids, "variant": var, "points": pts}
    with np.errstate(all="ignore"):
        ora = oracle_label(case, pts)
        _install(eng, case)
        cell = float(rng.choice([0.0, 0.0, 0.05, 0.4, 2.0]))
        got_b = _label(eng, case, pts, "brute")
        got_g = _label(eng, case, pts, "grid", cell_size=cell)
    assert_parity(got_b, ora, f"fuzz{seed}/{vn}/brute", require_bitwise=True)
    assert_parity(got_g, ora, f"fuzz{seed}/{vn}/grid(cell={cell})", require_bitwise=True)


# ---- full-size properties ---------------------------------------------------------------------------------

@pytest.mark.parametrize("n,m", [(1_000_000, 10_000), (10_000_000, 50_000)])
def test_full_size_properties(eng, n, m):
    """BASELINE.json configs[1] and configs[2] sizes: the grid answer equals exhaustive search on a random
    subset, is invariant under permutation of the points, and every reported distance is reproduced by
    re-evaluating the reported cylinder alone."""
    from treemorph_b200 import synth
    q = synth.random_qsm(m, seed=1)
    pts = synth.sample_points(q, n, seed=2)
    start, radius, length, unit, ids = synth.cylinder_arrays(q)
    case = {"start": start, "radius": radius, "length": length, "unit": unit, "ids": ids,
            "variant": _oracle.VARIANT_A, "points": pts}
    _install(eng, case)
    dev = eng.device
    dpts = torch.tensor(pts, device=dev)
    full = eng.label(dpts, api.VARIANT_A, mode="grid", want=("index", "id", "dist", "offset"))
    st = eng.stats()
    assert st["mode_used"] == 2 and st["points_grid"] + st["points_far"] + st["points_ring"] + st["points_tree"] + st["points_brute"] == n
    assert st["points_ring"] > 0 and st["points_tree"] == 0                                        # few stragglers: ring search
    rng = np.random.default_rng(3)
    sub = torch.tensor(rng.choice(n, 20_000, replace=False), device=dev)
    brute = eng.label(dpts[sub], api.VARIANT_A, mode="brute", want=("index", "id", "dist", "offset"))
    for k in ("index", "id", "dist", "offset"):
        a, b = full[k][sub], brute[k]
        same = (a == b) | (torch.isnan(a) & torch.isnan(b)) if a.dtype.is_floating_point else (a == b)
        assert bool(same.all()), f"{k}: grid != exhaustive on the subset"
    # oracle on a smaller subset (seconds on CPU)
    osub = sub[:4000].cpu().numpy()
    ora = oracle_label(case, pts[osub])
    got = {k: full[k][sub[:4000]].cpu().numpy() for k in ("index", "id", "dist", "offset")}
    assert_parity(got, ora, f"{n}x{m} vs oracle", require_bitwise=True)
    # permutation invariance
    perm = torch.randperm(n, device=dev)
    again = eng.label(dpts[perm], api.VARIANT_A, mode="grid", want=("index", "dist"))
    assert bool((again["index"] == full["index"][perm]).all())
    assert bool((again["dist"] == full["dist"][perm]).all())
    # idempotence of the winner: relabel against the single winning cylinder reproduces dist bit-for-bit
    pick = sub[:2000]
    win = full["index"][pick].long()
    d1 = torch.empty(len(pick), device=dev)
    for j in torch.unique(win)[:50]:
        rows = pick[win == j]
        eng.set_cylinders(torch.tensor(start[j:j + 1], device=dev), torch.tensor(radius[j:j + 1], device=dev),
                          torch.tensor(length[j:j + 1], device=dev), torch.tensor(unit[j:j + 1], device=dev))
        one = eng.label(dpts[rows], api.VARIANT_A, mode="brute", want=("dist",))
        assert bool((one["dist"] == full["dist"][rows]).all())


@pytest.mark.parametrize("vn", ["A", "B"])
def test_largest_table_of_the_sweep(eng, vn):
    """BASELINE.json configs[4]'s largest table (200k cylinders = a 40-tree plot) with 1M points, a tenth of them clutter
    metres away from any cylinder (tree search): equal to the exhaustive kernel on a subset and to the oracle on a smaller one."""
    from treemorph_b200 import synth
    m, n = 200_000, 1_000_000
    q = synth.random_qsm(m, seed=1003)
    pts = synth.sample_points(q, n, seed=2003)
    rng = np.random.default_rng(7)
    lo, hi = pts.min(0), pts.max(0)
    clutter = rng.choice(n, n // 10, replace=False)
    pts[clutter] = (lo + rng.random((len(clutter), 3)) * (hi - lo)).astype(np.float32)
    var = _oracle.VARIANTS[vn]
    start, radius, length, unit, ids = synth.cylinder_arrays(q, var.axis_eps)
    case = {"start": start, "radius": radius, "length": length, "unit": unit, "ids": ids, "variant": var, "points": pts}
    _install(eng, case)
    dev = eng.device
    dpts = torch.tensor(pts, device=dev)
    avar = api.VARIANTS[vn]
    full = eng.label(dpts, avar, mode="grid", want=("index", "id", "dist", "offset"))
    st = eng.stats()
    assert st["mode_used"] == 2 and st["points_tree"] > 50_000
    assert st["points_grid"] + st["points_far"] + st["points_ring"] + st["points_tree"] + st["points_brute"] == n
    sub_np = np.concatenate([rng.choice(n, 6_000, replace=False), clutter[:2_000]])
    sub = torch.tensor(sub_np, device=dev)
    brute = eng.label(dpts[sub], avar, mode="brute", want=("index", "id", "dist", "offset"))
    for k in ("index", "id", "dist", "offset"):
        a, b = full[k][sub], brute[k]
        same = (a == b) | (torch.isnan(a) & torch.isnan(b)) if a.dtype.is_floating_point else (a == b)
        assert bool(same.all()), f"{k}: grid != exhaustive on the subset"
    osub = np.concatenate([sub_np[:600], sub_np[-200:]])
    ora = oracle_label(case, pts[osub])
    got = {k: full[k][torch.tensor(osub, device=dev)].cpu().numpy() for k in ("index", "id", "dist", "offset")}
    assert_parity(got, ora, f"{n}x{m}/{vn} vs oracle", require_bitwise=True)


@pytest.mark.parametrize("m", [1, 2, 4, 5, 7, 8, 9, 16, 17, 31, 33, 64, 65, 129, 1000, 4097])
def test_tree_search_for_every_tree_shape(eng, m):
    """The bounding-volume hierarchy is built on the device from the table size alone (implicit split-in-the-middle tree,
    leaves of <= 4): root-is-a-leaf, odd splits and leaves at two depths must all give the exhaustive answer.  Half of the
    points lie far outside the voxel grid (tree search only), half are clutter inside it."""
    rng = np.random.default_rng(1000 + m)
    case = make_case(max(m, 1), 4, seed=200 + m, variant="A" if m % 2 else "B")
    lo = case["start"].min(0) - 1.0
    hi = case["start"].max(0) + 1.0
    n = 6000
    pts = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    pts[: n // 2] += (rng.normal(0, 1, (n // 2, 3)) * 40.0 + 60.0).astype(np.float32)
    case["points"] = pts
    _install(eng, case)
    with np.errstate(all="ignore"):
        got = _label(eng, case, pts, "grid", cell_size=0.25 if m > 4 else 0.0)
        if eng.stats()["mode_used"] == 2:
            assert eng.stats()["points_tree"] > 0
        want = _label(eng, case, pts, "brute")
    for k in ("index", "id", "dist", "offset"):
        assert_same_bits(got[k], want[k], f"m={m}: {k}")
    ora = oracle_label(case, pts[:: 10])
    assert_parity({k: v[:: 10] for k, v in got.items()}, ora, f"tree shapes m={m}", require_bitwise=True)


# ---- table cache of the kernel-level drop-in --------------------------------------------------------------

@pytest.mark.parametrize("vn", ["A", "B"])
def test_fresh_tables_of_the_same_shape_are_reinstalled(vn):
    """closest_cylinder_cuda_batch called the way cylinder_proximity_based_segmentation calls it (QSMFittingDepthFirst.py:
    1079-1081): brand-new cylinder tensors of the SAME shape for every call, the previous ones already freed — the caching
    allocator hands their addresses out again.  Every call must be answered against the table it was given."""
    mod = L if vn == "A" else P
    dev = torch.device("cuda")
    var = _oracle.VARIANTS[vn]
    pts = None
    for rep in range(6):
        case = make_case(300, 2000, seed=40 + rep, variant=vn)
        if pts is None:
            pts = case["points"]
        tabs = [torch.tensor(case[k], device=dev) for k in ("start", "radius", "length", "unit", "ids")]
        ids, dist, off = mod.closest_cylinder_cuda_batch(pts, *tabs, dev)
        # the same five tensor objects again: the table may be reused, the answer is the same
        ids2, dist2, off2 = mod.closest_cylinder_cuda_batch(pts, *tabs, dev)
        assert_same_bits(ids, ids2), assert_same_bits(dist, dist2)
        ora = _oracle.label(pts, case["start"], case["radius"], case["length"], case["unit"], case["ids"], var, norm_fma=True)
        assert_same_bits(ids, ora["id"], f"rep {rep}: ids")
        assert_same_bits(dist, ora["dist"], f"rep {rep}: distances")
        assert_same_bits(off, ora["offset"], f"rep {rep}: offsets")
        # in-place edit of an installed tensor: must be noticed as well
        tabs[1].mul_(1.5)
        ids3, dist3, _ = mod.closest_cylinder_cuda_batch(pts, *tabs, dev)
        ora3 = _oracle.label(pts, case["start"], case["radius"] * np.float32(1.5), case["length"], case["unit"], case["ids"], var,
                             norm_fma=True)
        assert_same_bits(ids3, ora3["id"], f"rep {rep}: ids after the in-place edit")
        assert_same_bits(dist3, ora3["dist"], f"rep {rep}: distances after the in-place edit")
        del tabs
        gc.collect()


# ---- BASELINE.json configs[3]: projection of model-corrected points, full size -----------------------------------

def test_config4_projection_5m_points_50k_cylinders():
    """5M PointTransformerV3-style points (N(0, 1 cm) residuals) x 50k cylinders through the Projection drop-in
    (Modules/Projection.py:117-144 signature; pageable float64 cloud in, (N,7) float64 out): equal to the oracle on a 20k-row
    sample, to the exhaustive kernel on another, xyz columns are the input's own float64 values."""
    from treemorph_b200 import synth
    n, m = 5_000_000, 50_000
    q = synth.random_qsm(m, seed=1)
    pts32 = synth.sample_points(q, n, seed=5, noise="model")
    cloud = pts32.astype(np.float64)
    out = P.generate_offset_cloud_cuda_batched(cloud, pd.DataFrame(q), torch.device("cuda"))
    assert out.shape == (n, 7) and out.dtype == np.float64
    assert np.array_equal(out[:, :3], cloud)
    rng = np.random.default_rng(9)
    rows = np.sort(rng.choice(n, 20_000, replace=False))
    ora = _oracle.label_cloud(cloud[rows], q, _oracle.VARIANT_B)
    assert_same_bits(out[rows], ora, "configs[3] vs oracle")
    # exhaustive kernel on the device for a second, larger sample
    rows2 = np.sort(rng.choice(n, 100_000, replace=False))
    e = api.get_engine(torch.device("cuda"))
    res = e.label(torch.tensor(pts32[rows2], device=e.device), api.VARIANT_B, mode="brute", want=("id", "offset"))
    assert np.array_equal(res["id"].cpu().numpy().astype(np.float64), out[rows2, 6])
    assert_same_bits(res["offset"].cpu().numpy().astype(np.float64), out[rows2, 3:6], "configs[3] vs exhaustive")


def test_full_size_properties_variant_b(eng):
    """configs[2] size in variant B (Projection.py arithmetic: atol 1e-3, guarded norms): grid == exhaustive == oracle on
    subsets, invariant under permutation."""
    from treemorph_b200 import synth
    n, m = 10_000_000, 50_000
    q = synth.random_qsm(m, seed=1)
    pts = synth.sample_points(q, n, seed=2)
    var = _oracle.VARIANT_B
    start, radius, length, unit, ids = synth.cylinder_arrays(q, var.axis_eps)
    case = {"start": start, "radius": radius, "length": length, "unit": unit, "ids": ids, "variant": var, "points": pts}
    _install(eng, case)
    dev = eng.device
    dpts = torch.tensor(pts, device=dev)
    full = eng.label(dpts, api.VARIANT_B, mode="grid", want=("index", "id", "dist", "offset"))
    rng = np.random.default_rng(4)
    sub = torch.tensor(rng.choice(n, 20_000, replace=False), device=dev)
    brute = eng.label(dpts[sub], api.VARIANT_B, mode="brute", want=("index", "id", "dist", "offset"))
    for k in ("index", "id", "dist", "offset"):
        a, b = full[k][sub], brute[k]
        same = (a == b) | (torch.isnan(a) & torch.isnan(b)) if a.dtype.is_floating_point else (a == b)
        assert bool(same.all()), f"{k}: grid != exhaustive on the subset"
    osub = sub[:4000].cpu().numpy()
    ora = oracle_label(case, pts[osub])
    got = {k: full[k][sub[:4000]].cpu().numpy() for k in ("index", "id", "dist", "offset")}
    assert_parity(got, ora, "10M x 50k variant B vs oracle", require_bitwise=True)
    perm = torch.randperm(n, device=dev)
    again = eng.label(dpts[perm], api.VARIANT_B, mode="grid", want=("index", "dist"))
    assert bool((again["index"] == full["index"][perm]).all())
    assert bool((again["dist"] == full["dist"][perm]).all())


# ---- round-2 entry points -----------------------------------------------------------------------------------

@pytest.mark.parametrize("f64", [False, True])
def test_wide_records_are_the_seven_columns_plus_the_tail(eng, f64):
    """tm_label_cloud_host_wide: the (N,11) rows the drivers save (LabelGenerationCuda.py:196-205) are the (N,7) record with
    the tail values appended, for float32 / float64 clouds with extra columns and a ragged last chunk."""
    case = make_case(700, 50_003, seed=81, variant="A")
    _install(eng, case)
    pts = case["points"]
    cloud = np.concatenate([pts.astype(np.float64) if f64 else pts, np.full((len(pts), 1), 3, np.float64 if f64 else pts.dtype)], axis=1)
    rec7 = eng.label_cloud_host(cloud, api.VARIANT_A, mode="grid")
    wide, dist = eng.label_cloud_host(cloud, api.VARIANT_A, mode="grid", tail=(1.0, 1.0, 1.0, 1.0), want_dist=True)
    assert wide.shape == (len(pts), 11) and wide.dtype == np.float64
    assert np.array_equal(wide[:, :7], rec7, equal_nan=True) and (wide[:, 7:] == 1.0).all()
    ora = oracle_label(case, pts)
    assert np.array_equal(dist, ora["dist"], equal_nan=True) and np.array_equal(wide[:, 6], ora["id"].astype(np.float64))
    odd = eng.label_cloud_host(cloud, api.VARIANT_A, mode="grid", tail=(0.5, -2.0))
    assert odd.shape == (len(pts), 9) and (odd[:, 7] == 0.5).all() and (odd[:, 8] == -2.0).all()
    with pytest.raises(ValueError):
        eng.label_cloud_host(cloud, api.VARIANT_A, tail=tuple(range(9)))             # rows of more than 15 doubles


def test_stats_of_the_estimate_path_and_host_probe(eng):
    from treemorph_b200 import synth
    q = synth.random_qsm(5000, seed=91)
    pts = synth.sample_points(q, 800_000, seed=92)                                     # above the direct path's limit: sorted path
    start, radius, length, unit, ids = synth.cylinder_arrays(q)
    case = {"start": start, "radius": radius, "length": length, "unit": unit, "ids": ids, "variant": _oracle.VARIANT_A, "points": pts}
    _install(eng, case)
    eng.label(torch.tensor(pts, device=eng.device), api.VARIANT_A, mode="grid")
    st = eng.stats()
    assert st["launches"] >= 10 and st["lane_ops_per_bound"] == 29
    assert st["bound_tests"] > 5 * len(pts) and 0 < st["points_slow"] < 0.3 * len(pts)
    assert st["pairs_evaluated"] < 1.0 * len(pts)                                      # far fewer exact evaluations than points
    assert st["points_grid"] + st["points_far"] + st["points_ring"] + st["points_tree"] + st["points_brute"] == len(pts)
    bw = eng.host_bandwidth()
    assert bw["threads"] >= 1 and 1e9 < bw["bytes_per_s"] < 5e12


def test_direct_and_sorted_paths_agree_in_fresh_processes(tmp_path):
    """TM_DIRECT is read once per process: label the same clutter-laden cloud with the sort-free path forced and forbidden
    in two subprocesses and compare every output bit."""
    import subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import sys, numpy as np, torch
        sys.path.insert(0, {root!r})
        from treemorph_b200 import api, synth
        q = synth.random_qsm(3000, seed=71)
        pts = synth.sample_points(q, 150_000, seed=72)
        rng = np.random.default_rng(73)
        pts[::7] += rng.normal(0, 2.0, size=pts[::7].shape).astype(np.float32)          # clutter metres away
        pts[5] = [np.nan, 0, 0]; pts[6] = [1e30, 1e30, 1e30]
        s, r, l, u, i = synth.cylinder_arrays(q)
        dev = torch.device("cuda", 0)
        e = api.Engine(dev)
        e.set_cylinders(*[torch.tensor(x, device=dev) for x in (s, r, l, u)], torch.tensor(i, device=dev))
        for vn in "AB":
            got = e.label(torch.tensor(pts, device=dev), api.VARIANTS[vn], mode="grid")
            np.savez(sys.argv[1] + vn + ".npz", **{{k: v.cpu().numpy() for k, v in got.items()}})
    """)
    outs = {}
    for flag in ("0", "1"):
        env = dict(os.environ, TM_DIRECT=flag)
        prefix = str(tmp_path / f"d{flag}_")
        subprocess.run([sys.executable, "-c", code, prefix], check=True, env=env, timeout=600)
        outs[flag] = {vn: dict(np.load(prefix + vn + ".npz")) for vn in "AB"}
    for vn in "AB":
        for k in ("index", "id", "dist", "offset"):
            a, b = outs["0"][vn][k], outs["1"][vn][k]
            assert np.array_equal(a.view(np.int32), b.view(np.int32)), f"variant {vn}: {k} differs between the two paths"


def test_table_caches_do_not_survive_an_engine_replacement():
    """The drop-ins skip re-installing a table they installed before.  The key carries the engine's serial number: a NEW engine
    whose install counter happens to match must not be taken for the one that holds the table."""
    from treemorph_b200 import dropin, synth
    from treemorph_b200.PreProcessing import LabelGenerationCuda as L
    dev = torch.device("cuda", torch.cuda.current_device())
    q1, q2 = synth.random_qsm(300, seed=61), synth.random_qsm(300, seed=62)
    cloud = synth.sample_points(q1, 4000, seed=63).astype(np.float64)
    df1, df2 = synth.qsm_dataframe(q1), synth.qsm_dataframe(q2)
    want1 = _oracle.label_cloud(cloud, q1, _oracle.VARIANT_A)
    assert np.array_equal(L.generate_offset_cloud_cuda_batched(cloud, df1, dev), want1, equal_nan=True)
    old = api.get_engine(dev)
    installs = old.installs
    old.close()                                                           # the process-wide engine is replaced ...
    fresh = api.get_engine(dev)
    assert fresh is not old and fresh.serial != old.serial
    while fresh.installs < installs:                                      # ... and brought to the same install count with ANOTHER table
        s, r, l, u, i = synth.cylinder_arrays(q2)
        fresh.set_cylinders(*[torch.tensor(x, device=dev) for x in (s, r, l, u)], torch.tensor(i, device=dev))
    assert fresh.installs == installs
    assert np.array_equal(L.generate_offset_cloud_cuda_batched(cloud, df1, dev), want1, equal_nan=True)
    want2 = _oracle.label_cloud(cloud, q2, _oracle.VARIANT_A)
    assert np.array_equal(L.generate_offset_cloud_cuda_batched(cloud, df2, dev), want2, equal_nan=True)


def test_dataframe_layouts_and_edits_reach_the_kernel():
    """The drop-in reads the QSM columns by name into one float32 block and re-installs the table only when that block
    differs bit for bit from the installed one: column order, extra columns and the frame's dtypes must not matter, an edited
    value (also a NaN that appears or a zero that changes sign) must."""
    from treemorph_b200 import synth
    from treemorph_b200.PreProcessing import LabelGenerationCuda as L
    dev = torch.device("cuda", torch.cuda.current_device())
    q = synth.random_qsm(400, seed=71)
    cloud = synth.sample_points(q, 5000, seed=72).astype(np.float64)
    df = synth.qsm_dataframe(q)
    want = _oracle.label_cloud(cloud, q, _oracle.VARIANT_A)
    eng = api.get_engine(dev)
    assert np.array_equal(L.generate_offset_cloud_cuda_batched(cloud, df, dev), want, equal_nan=True)
    installs = eng.installs
    # same values, another frame object: float32 columns in another order, extra columns, integer IDs as int64
    cols = list(df.columns)[::-1]
    other = df[cols].copy()
    for c in ("startX", "startY", "startZ", "endX", "endY", "endZ", "radius"):
        other[c] = other[c].astype(np.float32)
    other["ID"] = other["ID"].astype(np.int64)
    other["note"] = "x"
    other.insert(0, "volume", 1.0)
    assert np.array_equal(L.generate_offset_cloud_cuda_batched(cloud, other, dev), want, equal_nan=True)
    assert eng.installs == installs                                        # recognised by value: not installed again
    # one radius edited in place in the caller's frame
    edited = df.copy()
    edited.loc[edited.index[7], "radius"] *= 3.0
    q_edit = dict(q)
    q_edit["radius"] = np.array(q["radius"], dtype=np.float64, copy=True)
    q_edit["radius"][7] *= 3.0
    got = L.generate_offset_cloud_cuda_batched(cloud, edited, dev)
    assert eng.installs == installs + 1
    assert np.array_equal(got, _oracle.label_cloud(cloud, q_edit, _oracle.VARIANT_A), equal_nan=True)
    # a NaN row: installed once, recognised on the second call (NaN compares equal to itself bit for bit)
    broken = df.copy()
    broken.loc[broken.index[3], "endX"] = np.nan
    first = L.generate_offset_cloud_cuda_batched(cloud, broken, dev)
    count = eng.installs
    again = L.generate_offset_cloud_cuda_batched(cloud, broken.copy(), dev)
    assert eng.installs == count and np.array_equal(first, again, equal_nan=True)
    # and back to the original table
    assert np.array_equal(L.generate_offset_cloud_cuda_batched(cloud, df, dev), want, equal_nan=True)
