"""The claim the tile kernel's closed-form estimate rests on (csrc/tm_grid.cu: bound_pair), checked on
the CPU against the reference arithmetic:

    for every (point, regular cylinder) pair that bound_pair calls RELIABLE,   |sqrt(D) - dist_ref| <= S,

where dist_ref is the reference's fp32 distance (LabelGenerationCuda.py:36-84 / Projection.py:35-87 in mirror order) and
S = slack_floor + 4e-6 * max|coordinate| is the rounding allowance the kernel uses.  Unreliable pairs may be anything (they
are settled by reference-order evaluations).  The numpy mirror of the reference used here is itself pinned against the C
oracle (argmin and distance, bit for bit) in the same test.
"""
import numpy as np
import pytest

from oracle import oracle
from treemorph_b200 import synth


def mirror_all_pairs(pts, start, unit, length, radius, var):
    """dist_ref for every pair, fp32, one rounding per reference op."""
    f = np.float32
    p = pts[:, None, :].astype(f)
    s, u = start[None].astype(f), unit[None].astype(f)
    L, r = length.reshape(1, -1).astype(f), radius[None].astype(f)
    with np.errstate(all="ignore"):
        v = p - s
        t = ((v[..., 0] * u[..., 0] + v[..., 1] * u[..., 1]) + v[..., 2] * u[..., 2]).astype(f)
        tc = np.minimum(np.maximum(t, f(0)), L)
        q = s + tc[..., None] * u
        w = p - q
        d = ((w[..., 0] * u[..., 0] + w[..., 1] * u[..., 1]) + w[..., 2] * u[..., 2]).astype(f)
        perp = np.abs(d) <= f(var.perp_atol)
        rej = w - d[..., None] * u
        rho = np.sqrt((rej[..., 0] * rej[..., 0] + rej[..., 1] * rej[..., 1]) + rej[..., 2] * rej[..., 2]).astype(f)
        if var.norm_eps > 0:
            rho = np.where(rho < f(var.norm_eps), f(var.norm_eps), rho)
        n = rej / rho[..., None]
        h = n * r[..., None]
        nas, nae = q - h, q + h
        pn = p - nas
        pl = ((pn[..., 0] * n[..., 0] + pn[..., 1] * n[..., 1]) + pn[..., 2] * n[..., 2]).astype(f)
        plc = np.minimum(np.maximum(pl, f(0)), r + r)
        pona = nas + plc[..., None] * n
        fin = np.where(perp[..., None], nae, pona)
        e = p - fin
        return np.sqrt((e[..., 0] * e[..., 0] + e[..., 1] * e[..., 1]) + e[..., 2] * e[..., 2]).astype(f)


def estimate_all_pairs(pts, start, unit, length, radius, atol, S):
    """bound_pair in float64 (the kernel runs it in fp32 relative to the cylinder's start: errors far inside S)."""
    p = pts[:, None, :].astype(np.float64)
    s, u = start[None].astype(np.float64), unit[None].astype(np.float64)
    L, r = length.reshape(1, -1).astype(np.float64), radius[None].astype(np.float64)
    v = p - s
    t = (v * u).sum(-1)
    tc = np.clip(t, 0.0, L)
    d = t - tc
    rej = v - t[..., None] * u
    rho = np.sqrt((rej ** 2).sum(-1))
    a = rho - r
    ad = np.abs(d)
    beyond = ad > atol + S
    undecided = ~beyond & (ad >= atol - S) if atol > 2 * S else ~beyond
    asel = np.where(beyond, np.maximum(a, 0.0), a)
    est = np.sqrt(asel ** 2 + d ** 2)
    unreliable = (undecided & (a < S)) | (rho < S)
    return est, unreliable


def _cases():
    # a synthetic tree with its noisy surface cloud, interior / on-axis points added
    q = synth.random_qsm(400, seed=11)
    pts = synth.sample_points(q, 1500, seed=12)
    s = np.stack([q["startX"], q["startY"], q["startZ"]], 1).astype(np.float32)
    e = np.stack([q["endX"], q["endY"], q["endZ"]], 1).astype(np.float32)
    pts[::9] = (0.5 * (s[:len(pts[::9])] + e[:len(pts[::9])])).astype(np.float32)              # on the axes, inside the slabs
    pts[1::9] = (s[:len(pts[1::9])] + 1.2 * (e[:len(pts[1::9])] - s[:len(pts[1::9])])).astype(np.float32)   # on the axis lines, beyond the caps
    yield "tree", s, e, np.asarray(q["radius"], np.float32), pts, 0.25
    # the same tree far from the origin (coarser fp32 grid) and at millimetre scale
    shift = np.array([310.0, -120.0, 45.0], np.float32)
    yield "tree+shift", s + shift, e + shift, np.asarray(q["radius"], np.float32), pts + shift, 0.25
    yield "tree*0.01", s * np.float32(0.01), e * np.float32(0.01), np.asarray(q["radius"], np.float32) * np.float32(0.01), pts * np.float32(0.01), 0.0025


@pytest.mark.parametrize("vn", ["A", "B"])
def test_reliable_estimates_are_within_the_rounding_allowance(vn):
    var = oracle.VARIANTS[vn]
    for name, s, e, r, pts, h in _cases():
        with np.errstate(all="ignore"):
            length, unit = oracle.prepare(s, e, var)
        regular = np.isfinite(unit).all(1) & (np.abs((unit.astype(np.float64) ** 2).sum(1) - 1.0) <= 4e-6)
        ref = mirror_all_pairs(pts, s, unit, length, r, var)
        # pin the mirror: its argmin / distance are the C oracle's
        ora = oracle.label(pts, s, r, length, unit, np.arange(len(s), dtype=np.int32), var)
        with np.errstate(invalid="ignore"):
            key = np.where(np.isnan(ref), -1.0, ref.astype(np.float64))
        assert (np.argmin(key, axis=1) == ora["index"]).all(), f"{name}: numpy mirror and C oracle disagree on the argmin"
        assert np.array_equal(ref[np.arange(len(pts)), ora["index"]], ora["dist"], equal_nan=True)
        maxabs = float(max(np.abs(s).max(), np.abs(e).max())) + float(r.max()) + 6 * h
        S = min(1e-4, 4e-4 * h) + 4e-6 * maxabs                      # tm_grid.cu: slack_for
        est, unreliable = estimate_all_pairs(pts, s, unit, length, r, var.perp_atol, S)
        ok = regular[None, :] & ~unreliable
        with np.errstate(invalid="ignore"):
            err = np.abs(est - ref.astype(np.float64))
        assert not np.isnan(ref[ok]).any(), f"{name}/{vn}: a reliable pair is NaN in the reference"
        worst = float(err[ok].max())
        assert worst <= S, f"{name}/{vn}: reliable estimate off by {worst:.3e} m, allowance {S:.3e} m"
        assert worst <= 0.25 * S, f"{name}/{vn}: margin thinner than expected ({worst:.3e} of {S:.3e})"
        # the unreliable set is small on a surface cloud and contains every NaN of the reference
        assert unreliable[:, regular].mean() < 0.05
        assert (unreliable | ~regular[None, :])[np.isnan(ref)].all()
