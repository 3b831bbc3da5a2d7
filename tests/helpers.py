"""Shared helpers of the parity tests: seeded inputs, oracle calls, comparison with the near-tie rule."""
from __future__ import annotations

import numpy as np

from oracle import oracle
from treemorph_b200 import synth

ABS_TOL_M = 1e-5          # north_star: offsets and distances within 1e-5 m absolute
NEAR_TIE_REL = 1e-6       # north_star: index mismatches allowed only where the top-2 distances differ by < 1e-6 relative


def make_case(m, n, seed=1, noise="lognormal", variant="A", id_offset=0):
    var = oracle.VARIANTS[variant]
    q = synth.random_qsm(m, seed=seed, id_offset=id_offset)
    pts = synth.sample_points(q, n, seed=seed + 1, noise=noise)
    start, radius, length, unit, ids = synth.cylinder_arrays(q, var.axis_eps)
    return {"qsm": q, "points": pts, "start": start, "radius": radius, "length": length, "unit": unit, "ids": ids,
            "variant": var}


def oracle_label(case, points=None, **kw):
    pts = case["points"] if points is None else points
    return oracle.label(pts, case["start"], case["radius"], case["length"], case["unit"], case["ids"],
                        case["variant"], norm_fma=kw.pop("norm_fma", False), **kw)


def same_or_nan(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return (a == b) | (np.isnan(a) & np.isnan(b))


def assert_parity(got: dict, ora: dict, what: str = "", require_bitwise: bool = False):
    """got: dict of numpy arrays (index, id, dist, offset).  ora: oracle.label output.

    Indices must be equal except at documented near-ties (oracle top-2 gap < 1e-6 relative);
    distances / offsets within 1e-5 m (NaN patterns must agree)."""
    idx_ok = got["index"] == ora["index"]
    bad = np.flatnonzero(~idx_ok)
    if bad.size:
        d, s = ora["dist"][bad].astype(np.float64), ora["second"][bad].astype(np.float64)
        near_tie = (s - d) <= NEAR_TIE_REL * np.maximum(d, 1e-30)
        mine = got["dist"][bad].astype(np.float64)
        within = np.abs(mine - d) <= NEAR_TIE_REL * np.maximum(d, 1e-30) * 2
        assert (near_tie & within).all(), (
            f"{what}: {np.count_nonzero(~(near_tie & within))} index mismatches outside the near-tie window "
            f"(first rows {bad[~(near_tie & within)][:5]})")
    if "id" in got:
        assert (got["id"][idx_ok] == ora["id"][idx_ok]).all(), f"{what}: id differs where index agrees"
    nan_g, nan_o = np.isnan(got["dist"]), np.isnan(ora["dist"])
    assert (nan_g == nan_o).all(), f"{what}: NaN pattern of distances differs"
    fin = ~nan_o
    with np.errstate(invalid="ignore"):
        gd, od = got["dist"][fin].astype(np.float64), ora["dist"][fin].astype(np.float64)
        dd = np.where(gd == od, 0.0, np.abs(gd - od))             # equal infinities count as equal
    assert dd.size == 0 or dd.max() <= ABS_TOL_M, f"{what}: distance differs by {dd.max():.3g} m"
    ok_rows = idx_ok & fin
    with np.errstate(invalid="ignore"):
        go, oo = got["offset"][ok_rows].astype(np.float64), ora["offset"][ok_rows].astype(np.float64)
        do = np.where(go == oo, 0.0, np.abs(go - oo))
    assert (np.isnan(go) == np.isnan(oo)).all(), f"{what}: NaN pattern of offsets differs"
    do = do[~np.isnan(do)]
    assert do.size == 0 or do.max() <= ABS_TOL_M, f"{what}: offset differs by {do.max():.3g} m"
    if require_bitwise:
        assert bad.size == 0, f"{what}: {bad.size} index mismatches"
        assert same_or_nan(got["dist"], ora["dist"]).all(), f"{what}: distances not bit-identical"
        assert same_or_nan(got["offset"], ora["offset"]).all(), f"{what}: offsets not bit-identical"
    return int(bad.size)
