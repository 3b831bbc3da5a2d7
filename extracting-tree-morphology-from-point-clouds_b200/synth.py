"""Synthetic QSMs and point clouds (there is no network, so no real trees).

The distributions follow SURVEY.md §8(d):

* ``random_qsm``   – a stochastic branching random walk in metres (cylinder table with the
  reference's CSV schema ``startX..endZ, radius, ID``; see a8 in SURVEY.md §8 and
  ``PreProcessing/LabelGenerationCuda.py:117-120`` for the columns the hot path reads).
* ``sample_points`` – NoiseDataGeneration-style surface samples: cylinder chosen in proportion
  to its lateral area, uniform angle / height, radial distance ``r + lognormal(-3, 0.85)``
  (``PreProcessing/NoiseDataGeneration.py:56-75``), rotated into the cylinder frame, fp32.

Only numpy is used here; pandas is imported lazily for ``qsm_dataframe``.
"""
from __future__ import annotations

import numpy as np

TREE_CYLINDERS = 5000      # one synthetic tree; larger tables are plots of several trees
PLOT_SPACING_M = 5.0


def _one_tree(m: int, rng: np.random.Generator) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Branching random walk: returns (start (m,3), end (m,3), radius (m,)) in float64."""
    start = np.empty((m, 3))
    end = np.empty((m, 3))
    radius = np.empty(m)
    tip_pos = [np.zeros(3)]
    tip_dir = [np.array([0.0, 0.0, 1.0])]
    tip_rad = [0.25]
    # draw all randomness up front (keeps the loop cheap and the stream layout fixed)
    pick = rng.random(m)
    kick = rng.normal(0.0, 0.15, size=(m, 3))
    seglen = rng.uniform(0.05, 0.30, size=m)
    taper = rng.uniform(0.9, 1.0, size=m)
    fork = rng.random(m) < 0.15
    fork_kick = rng.normal(0.0, 0.8, size=(m, 3))
    for i in range(m):
        t = int(pick[i] * len(tip_pos))
        d = tip_dir[t] + kick[i]
        d /= np.linalg.norm(d)
        p0 = tip_pos[t]
        p1 = p0 + seglen[i] * d
        start[i] = p0
        end[i] = p1
        radius[i] = tip_rad[t]
        r_next = max(tip_rad[t] * taper[i], 0.003)
        tip_pos[t] = p1
        tip_dir[t] = d
        tip_rad[t] = r_next
        if fork[i]:
            sd = d + fork_kick[i]
            n = np.linalg.norm(sd)
            sd = sd / n if n > 1e-9 else d
            tip_pos.append(p1.copy())
            tip_dir.append(sd)
            tip_rad.append(max(r_next * 0.6, 0.003))
    return start, end, radius


def random_qsm(m: int, seed: int = 1, id_offset: int = 0) -> dict[str, np.ndarray]:
    """Cylinder table with ``m`` rows as a dict of float64 columns (+ int64 ``ID``).

    ``m <= TREE_CYLINDERS`` is a single tree at the origin; larger tables are
    ``ceil(m / TREE_CYLINDERS)`` independent trees on a jittered 5 m grid (a "plot").
    """
    rng = np.random.default_rng(seed)
    n_trees = max(1, -(-m // TREE_CYLINDERS))
    per = [m // n_trees + (1 if t < m % n_trees else 0) for t in range(n_trees)]
    side = int(np.ceil(np.sqrt(n_trees)))
    starts, ends, rads = [], [], []
    for t, mt in enumerate(per):
        s, e, r = _one_tree(mt, rng)
        if n_trees > 1:
            shift = np.array([(t % side) * PLOT_SPACING_M, (t // side) * PLOT_SPACING_M, 0.0])
            shift[:2] += rng.uniform(-1.0, 1.0, size=2)
            s = s + shift
            e = e + shift
        starts.append(s)
        ends.append(e)
        rads.append(r)
    s = np.concatenate(starts)
    e = np.concatenate(ends)
    r = np.concatenate(rads)
    return {
        "startX": s[:, 0].copy(), "startY": s[:, 1].copy(), "startZ": s[:, 2].copy(),
        "endX": e[:, 0].copy(), "endY": e[:, 1].copy(), "endZ": e[:, 2].copy(),
        "radius": r, "ID": np.arange(m, dtype=np.int64) + id_offset,
    }


def qsm_dataframe(qsm: dict[str, np.ndarray]):
    """The same table as the pandas DataFrame ``pd.read_csv`` would yield for a QSM file."""
    import pandas as pd
    return pd.DataFrame(qsm)


def _frames(axis_unit: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Two unit vectors spanning the plane orthogonal to each axis."""
    helper = np.where(np.abs(axis_unit[:, 2:3]) < 0.9, [[0.0, 0.0, 1.0]], [[1.0, 0.0, 0.0]])
    e1 = np.cross(axis_unit, helper)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(axis_unit, e1)
    return e1, e2


def sample_points(qsm: dict[str, np.ndarray], n: int, seed: int = 2, noise: str = "lognormal",
                  chunk: int = 4_000_000) -> np.ndarray:
    """``n`` noisy surface samples of the QSM as float32 ``(n, 3)``.

    ``noise="lognormal"`` is the NoiseDataGeneration residual ``lognormal(-3, 0.85)``;
    ``noise="model"`` is the PTv3-corrected-cloud residual ``N(0, 0.01^2)`` used by the
    projection configuration (BASELINE.json configs[3]).
    """
    rng = np.random.default_rng(seed)
    s = np.stack([qsm["startX"], qsm["startY"], qsm["startZ"]], axis=1)
    e = np.stack([qsm["endX"], qsm["endY"], qsm["endZ"]], axis=1)
    r = np.asarray(qsm["radius"], dtype=np.float64)
    axis = e - s
    length = np.linalg.norm(axis, axis=1)
    unit = axis / np.maximum(length, 1e-12)[:, None]
    e1, e2 = _frames(unit)
    w = r * length
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    out = np.empty((n, 3), dtype=np.float32)
    for lo in range(0, n, chunk):
        k = min(chunk, n - lo)
        c = np.minimum(np.searchsorted(cdf, rng.random(k)), len(r) - 1)
        theta = rng.uniform(0.0, 2.0 * np.pi, k)
        z = rng.uniform(0.0, 1.0, k) * length[c]
        if noise == "lognormal":
            rad = r[c] + rng.lognormal(-3.0, 0.85, k)
        elif noise == "model":
            rad = r[c] + rng.normal(0.0, 0.01, k)
        else:
            raise ValueError(f"unknown noise model {noise!r}")
        p = (s[c] + z[:, None] * unit[c]
             + (rad * np.cos(theta))[:, None] * e1[c]
             + (rad * np.sin(theta))[:, None] * e2[c])
        out[lo:lo + k] = p.astype(np.float32)
    return out


def cylinder_arrays(qsm: dict[str, np.ndarray], guard_eps: float = 0.0):
    """fp32 kernel inputs built the way ``generate_offset_cloud_cuda_batched`` builds them
    (``LabelGenerationCuda.py:117-123``; ``Projection.py:121-132`` with ``guard_eps=1e-8``),
    in numpy, row-major.  Returns (start, radius, axis_length (M,1), axis_unit, ids int32).
    """
    start = np.stack([qsm["startX"], qsm["startY"], qsm["startZ"]], axis=1).astype(np.float32)
    end = np.stack([qsm["endX"], qsm["endY"], qsm["endZ"]], axis=1).astype(np.float32)
    radius = np.asarray(qsm["radius"]).astype(np.float32)
    ids = np.asarray(qsm["ID"]).astype(np.int32)
    axis = end - start
    sq = axis * axis
    length = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2]).astype(np.float32)[:, None]
    div = length.copy()
    if guard_eps > 0.0:
        div[div < np.float32(guard_eps)] = np.float32(guard_eps)
    with np.errstate(divide="ignore", invalid="ignore"):
        unit = (axis / div).astype(np.float32)
    return start, radius, length, unit, ids
