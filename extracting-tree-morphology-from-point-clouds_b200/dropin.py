"""Shared body of the two drop-in modules.

``PreProcessing/LabelGenerationCuda.py`` and ``Modules/Projection.py`` of the reference carry two copies
of the same kernel that differ in three constants; here one implementation is
parameterised by ``api.Variant`` and both modules forward to it.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import api

QSM_COLUMNS = ("startX", "startY", "startZ", "endX", "endY", "endZ", "radius", "ID")

# device index -> (the caller's cylinder tensor OBJECTS, their _version counters, the engine's install counter).
# The objects are held strongly: an address can be recycled by the caching allocator the moment its tensor dies, so
# only "the very same live tensor objects, unmodified, and nobody installed another table since" skips the install.
_table_key: dict[int, tuple] = {}
# device index -> ((variant, norm_fma, engine serial, engine install counter), (M,7) float32 table, int32 IDs) of the last
# DataFrame install
_frame_key: dict[int, tuple] = {}


def _cuda_device(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"closest_cylinder_cuda_batch needs a CUDA device, got {dev}: "
                           "the B200 path has no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _same_table(dev_index: int, eng, given: tuple) -> bool:
    held = _table_key.get(dev_index)
    if held is None:
        return False
    objs, versions, installs = held
    return (installs == (eng.serial, eng.installs) and len(objs) == len(given)
            and all(a is b for a, b in zip(objs, given))
            and all(t._version == v for t, v in zip(given, versions)))


def closest_cylinder(points, start, radius, axis_length, axis_unit, IDs, device, variant: api.Variant,
                     move_points_to_mantle: bool = True):
    """``closest_cylinder_cuda_batch`` (LabelGenerationCuda.py:20-111 / Projection.py:19-115).

    points: (N,3) array-like or tensor; cylinder arguments: torch tensors (any strides) or array-likes.
    Returns host numpy ``(ids int32 (N,), distances float32 (N,), offsets float32 (N,3))``.
    """
    dev = _cuda_device(device)
    eng = api.get_engine(dev)
    given = (start, radius, axis_length, axis_unit, IDs)
    cyl = []
    for t, dt in zip(given, (torch.float32, torch.float32, torch.float32, torch.float32, torch.int32)):
        t = torch.as_tensor(t)
        if t.device != dev or t.dtype != dt:
            t = t.to(device=dev, dtype=dt)
        cyl.append(t)
    start_t, radius_t, length_t, unit_t, ids_t = cyl
    # the install is skipped only for the caller's own tensor objects (label_clouds-style loops pass the same five
    # tensors for every batch); anything converted on the way in, or not a tensor, is installed again
    reusable = all(isinstance(g, torch.Tensor) and g is c for g, c in zip(given, cyl))
    if not (reusable and _same_table(dev.index, eng, given)):
        eng.set_cylinders(start_t, radius_t, length_t, unit_t, ids_t)
        if reusable:
            _table_key[dev.index] = (given, tuple(t._version for t in given), (eng.serial, eng.installs))
        else:
            _table_key.pop(dev.index, None)
    m = start_t.shape[0]
    # ATen rounds norm() differently for a contiguous xyz axis (C-ordered tensors, or M == 1) than for
    # the Fortran-ordered tensors the DataFrame path produces; mirror whichever the caller's layout implies.
    norm_fma = m <= 1 or start_t.stride(1) == 1
    pts = torch.as_tensor(np.asarray(points) if not isinstance(points, torch.Tensor) else points)
    pts = pts.to(device=dev, dtype=torch.float32)            # torch.tensor(points, dtype=float32, device=device)
    res = eng.label(pts, variant, move_to_mantle=move_points_to_mantle, norm_fma=norm_fma,
                    want=("id", "dist", "offset"))
    return res["id"].cpu().numpy(), res["dist"].cpu().numpy(), res["offset"].cpu().numpy()


def offset_cloud(cloud, cylinders, device, variant: api.Variant, masterBar=None, batch_size=1024, out=None, tail=None):
    """``generate_offset_cloud_cuda_batched`` (LabelGenerationCuda.py:113-135 / Projection.py:117-144).

    cloud: (N, >=3) array; cylinders: DataFrame with the QSM columns.  Returns float64 (N,7)
    ``[x, y, z, ox, oy, oz, ID]``.  ``batch_size`` is accepted for signature compatibility; results do not
    depend on it (each point is independent), so the whole cloud is processed in pipelined chunks.
    """
    del masterBar, batch_size
    dev = _cuda_device(device)
    eng = api.get_engine(dev)
    # one (7, M) float32 block, filled column by column (DataFrame[[...]] re-indexes and copies through a block manager, and
    # a row-major (M,7) block costs strided writes: 0.5 ms instead of 1.7 at 50k cylinders)
    m = len(cylinders)
    table = np.empty((7, m), dtype=np.float32)
    for k, name in enumerate(QSM_COLUMNS[:7]):
        table[k] = cylinders[name].to_numpy()
    ids_np = np.asarray(cylinders["ID"].to_numpy()).astype(np.int32)
    norm_fma = m <= 1                        # DataFrame tensors are Fortran-ordered in the reference (strided norm)
    # the same QSM again (augmented clouds of one tree, repeated calls): the table and its voxel index are still installed.
    # Decided on the VALUES, bit for bit (a 1.6 MB comparison at 50k cylinders: 0.2 ms), never on object identity.
    held = _frame_key.get(dev.index)
    if not (held is not None and held[0] == (variant, norm_fma, eng.serial, eng.installs) and held[1].shape == table.shape
            and np.array_equal(held[1].view(np.uint32), table.view(np.uint32)) and np.array_equal(held[2], ids_np)):
        start = torch.as_tensor(table[0:3].T, device=dev)        # (M,3) views with strides (1, M): the reference's own layout
        end = torch.as_tensor(table[3:6].T, device=dev)
        radius = torch.as_tensor(table[6], device=dev)
        ids = torch.as_tensor(ids_np, device=dev)
        length, unit = eng.prepare(start, end, variant, norm_fma=norm_fma)
        eng.set_cylinders(start, radius, length, unit, ids)
        _frame_key[dev.index] = ((variant, norm_fma, eng.serial, eng.installs), table, ids_np)
        _table_key.pop(dev.index, None)
    cloud = np.asarray(cloud)
    if cloud.ndim != 2 or cloud.shape[1] < 3:
        raise ValueError(f"cloud must have shape (N, >=3), got {cloud.shape}")
    if len(cloud) == 0:
        return np.zeros((0, 7 + (len(tail) if tail is not None else 0)))
    return eng.label_cloud_host(cloud, variant, norm_fma=norm_fma, out=out, tail=tail)


def offset_cloud_to_npy(path, cloud, cylinders, device, variant: api.Variant, tail=(1.0, 1.0, 1.0, 1.0)):
    """``np.save(path, np.concatenate([generate_offset_cloud_cuda_batched(...), ones((N,4))], axis=1))`` in one pass
    (LabelGenerationCuda.py:196-205, Projection.py:416-438 without features): the host workers of the library write the
    (N,11) rows once, into a page-locked array that is reused from call to call (no first-touch page faults, no
    ``np.concatenate`` copy), and the file is written from it by a few threads with ``pwrite`` — the same bytes ``np.save``
    produces (format 1.0 header, C-ordered little-endian float64)."""
    n = len(cloud)
    width = 7 + len(tail)
    if n == 0:
        np.save(path, np.zeros((0, width)))
        return
    rows = api.Engine._new_records(n, width)
    offset_cloud(cloud, cylinders, device, variant, out=rows, tail=tail)
    _write_npy(path, rows)


def _write_npy(path, array: np.ndarray, threads: int = 4) -> None:
    """np.save for a C-contiguous array, with the payload copied into the page cache by several threads."""
    import io
    from concurrent.futures import ThreadPoolExecutor
    head = io.BytesIO()
    np.lib.format.write_array_header_1_0(head, np.lib.format.header_data_from_array_1_0(array))
    header = head.getvalue()
    payload = memoryview(np.ascontiguousarray(array).reshape(-1).view(np.uint8))
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o666)
    try:
        os.write(fd, header)
        size = len(payload)
        step = max(8 << 20, -(-size // threads))
        spans = [(lo, min(size, lo + step)) for lo in range(0, size, step)] or [(0, 0)]

        def put(span):
            lo, hi = span
            while lo < hi:
                lo += os.pwrite(fd, payload[lo:hi], len(header) + lo)
        if len(spans) > 1:
            with ThreadPoolExecutor(len(spans)) as pool:
                list(pool.map(put, spans))
        else:
            put(spans[0])
    finally:
        os.close(fd)
