"""TEST INFRASTRUCTURE — the reference's tensor-op chain restated in torch, runnable on ``cuda`` and ``cpu``.

Why it exists: the C oracle mirrors the roundings of ATen's *CPU* kernels (probed in the build container).  The
reference's normal deployment is ``device=cuda`` (``Modules/Utils.py:146-158``, ``LabelGenerationCuda.py:33``), and
ATen's CUDA ``sum`` / ``norm`` reductions over the strided xyz axis need not round like their CPU counterparts.
``/root/reference`` does not exist on the GPU box, so the reference itself cannot run there; this module issues the
same ATen calls in the same order (one torch op per row of SURVEY.md A.1) so that the box can answer "what does the
reference compute on a B200" — it is the comparator of ``scripts/reference_cuda_check.py`` and of
``tests/test_reference_cuda.py`` and is never imported by the product path.

Each step cites the reference line it stands for: A = ``PreProcessing/LabelGenerationCuda.py``,
B = ``Modules/Projection.py``.
"""
from __future__ import annotations

import numpy as np
import torch


def table_tensors(qsm: dict, device, axis_eps: float = 0.0, fortran: bool = True):
    """Cylinder tensors as ``generate_offset_cloud_cuda_batched`` builds them (A:117-123 / B:121-132).

    ``fortran=True`` reproduces the layout of ``DataFrame[[...]].values`` (column-major, strides (1, M)); the arrays are
    created on the host with those strides and moved to ``device`` by ``torch.tensor`` exactly as the reference does.
    """
    def block(names):
        cols = np.stack([np.asarray(qsm[k], dtype=np.float64) for k in names], axis=1)
        return np.asfortranarray(cols) if fortran else np.ascontiguousarray(cols)
    start = torch.tensor(block(("startX", "startY", "startZ")), dtype=torch.float32, device=device)      # A:117
    end = torch.tensor(block(("endX", "endY", "endZ")), dtype=torch.float32, device=device)              # A:118
    radius = torch.tensor(np.asarray(qsm["radius"], dtype=np.float64), dtype=torch.float32, device=device)
    ids = torch.tensor(np.asarray(qsm["ID"]), dtype=torch.int32, device=device)
    axis = end - start                                                                                    # A:121
    length = torch.norm(axis, dim=1, keepdim=True)                                                        # A:122
    if axis_eps > 0:                                                                                      # B:129-132
        safe = length.clone()
        safe[safe < axis_eps] = axis_eps
        unit = axis / safe
    else:
        unit = axis / length                                                                              # A:123
    return start, radius, length, unit, ids


def closest(points, start, radius, length, unit, ids, device, perp_atol: float, norm_eps: float,
            want_matrix: bool = False):
    """One batch of ``closest_cylinder_cuda_batch`` → (ids, distances, offsets, indices) as host numpy."""
    p = torch.tensor(points, dtype=torch.float32, device=device)[:, None, :]                              # A:33
    s, u = start[None, :, :], unit[None, :, :]
    r = radius.view(1, -1, 1)
    v = p - s                                                                                             # A:36
    t = torch.sum(v * u, dim=2, keepdim=True)                                                             # A:39
    tc = torch.clamp(t, torch.zeros_like(t), length[None, :, :])                                          # A:42-43
    q = s + tc * u                                                                                        # A:44
    w = p - q                                                                                             # A:47
    d = torch.sum(w * u, dim=2)                                                                           # A:50
    perp = torch.isclose(d, torch.tensor(0.0, device=device), atol=perp_atol)                             # A:51 / B:50
    rej = w - d[..., None] * u                                                                            # A:54-55
    rho = torch.norm(rej, dim=2, keepdim=True)                                                            # A:58
    if norm_eps > 0:                                                                                      # B:60-62
        rho = rho.clone()
        rho[rho < norm_eps] = norm_eps
    nh = rej / rho                                                                                        # A:60
    sc = nh * (2 * r)                                                                                     # A:63
    nas = q - 0.5 * sc                                                                                    # A:66
    nae = q + 0.5 * sc                                                                                    # A:67
    pl = torch.sum((p - nas) * nh, dim=2, keepdim=True)                                                   # A:70
    plc = torch.clamp(pl, torch.zeros_like(pl), 2 * r)                                                    # A:73-74
    pona = nas + plc * nh                                                                                 # A:75
    surf = q + rej / rho * r                                                                              # A:78
    fin = torch.where(perp[..., None], surf, pona)                                                        # A:81
    dist = torch.norm(p - fin, dim=2)                                                                     # A:84
    j = torch.argmin(dist, dim=1)                                                                         # A:87
    rows = torch.arange(dist.shape[0], device=device)
    best = dist[rows, j]                                                                                  # A:88
    ds = torch.norm(pona - nas, dim=2, keepdim=True)                                                      # A:92
    de = torch.norm(pona - nae, dim=2, keepdim=True)                                                      # A:93
    face = torch.where(ds < de, nas, nae)                                                                 # A:96-97
    mantle = torch.where(perp[..., None], surf, face)                                                     # A:100
    off = mantle[rows, j] - p[:, 0, :]                                                                    # A:103-106
    out = (ids[j].cpu().numpy(), best.cpu().numpy(), off.cpu().numpy(), j.cpu().numpy().astype(np.int32))
    if want_matrix:
        return out + (dist.cpu().numpy(),)
    return out


def label(points: np.ndarray, qsm: dict, device, perp_atol: float, norm_eps: float, axis_eps: float,
          batch_size: int = 1024, fortran: bool = True):
    """The batch loop of ``generate_offset_cloud_cuda_batched`` (A:126-133): dict(index, id, dist, offset)."""
    tabs = table_tensors(qsm, device, axis_eps, fortran)
    n = len(points)
    res = {"index": np.empty(n, np.int32), "id": np.empty(n, np.int32), "dist": np.empty(n, np.float32),
           "offset": np.empty((n, 3), np.float32)}
    for lo in range(0, n, batch_size):
        ids, dist, off, idx = closest(points[lo:lo + batch_size, :3], *tabs, device, perp_atol, norm_eps)
        res["id"][lo:lo + batch_size] = ids
        res["dist"][lo:lo + batch_size] = dist
        res["offset"][lo:lo + batch_size] = off
        res["index"][lo:lo + batch_size] = idx
    return res
