"""Multi-GPU: one process per GPU, points sharded, cylinder table replicated.

The nearest-cylinder search has no cross-point dependency: every point needs the whole cylinder table and
nothing else.  So the only communication is ONE broadcast of the packed table from the
rank that read the QSM (NCCL over NVLink on GPUs, gloo in the CPU tests); after that each rank labels its own
contiguous slice of the cloud and there is no data-path collective.  ``gather_records`` is optional plumbing
for callers that want the whole ``(N,7)`` record on one rank.

torch.distributed is used as the launcher/communicator only (``torchrun`` sets RANK / WORLD_SIZE / MASTER_*).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

TABLE_COLUMNS = 9       # start xyz | axis_unit xyz | axis_length | radius | ID (int32 bits carried in a float32 lane)


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of an n-point cloud owned by `rank`; sizes differ by at most one point."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_table(start, radius, axis_length, axis_unit, ids) -> torch.Tensor:
    """(M,9) float32; the int32 IDs travel bit-for-bit in the last column."""
    start = torch.as_tensor(start, dtype=torch.float32).reshape(-1, 3)
    m = start.shape[0]
    table = torch.empty((m, TABLE_COLUMNS), dtype=torch.float32, device=start.device)
    table[:, 0:3] = start
    table[:, 3:6] = torch.as_tensor(axis_unit, dtype=torch.float32, device=start.device).reshape(-1, 3)
    table[:, 6] = torch.as_tensor(axis_length, dtype=torch.float32, device=start.device).reshape(-1)
    table[:, 7] = torch.as_tensor(radius, dtype=torch.float32, device=start.device).reshape(-1)
    table[:, 8] = torch.as_tensor(ids, dtype=torch.int32, device=start.device).reshape(-1).view(torch.float32)
    return table


def unpack_table(table: torch.Tensor):
    """→ start (M,3), radius (M,), axis_length (M,1), axis_unit (M,3), ids int32 (M,) — views, no copies."""
    return (table[:, 0:3], table[:, 7], table[:, 6:7], table[:, 3:6], table[:, 8].contiguous().view(torch.int32))


def broadcast_table(table: torch.Tensor | None, device, src: int = 0) -> torch.Tensor:
    """Replicate the packed table from `src` to every rank (one size message + one payload message)."""
    rank, size = world()
    if size == 1:
        assert table is not None
        return table.to(device)
    count = torch.tensor([table.shape[0] if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(count, src=src)
    m = int(count.item())
    if rank == src:
        buf = table.to(device=device, dtype=torch.float32).contiguous()
    else:
        buf = torch.empty((m, TABLE_COLUMNS), dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src)
    return buf


def init_engine_comm(engine) -> None:
    """Give `engine` its own NCCL communicator through the C ABI (tm_comm_init_rank): rank 0 makes the unique id, torch's
    process group only ships its 128 bytes.  After this ``engine.broadcast_cylinders`` replicates and installs a table
    without torch.distributed on the path."""
    rank, size = world()
    if size == 1:
        return
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.comm_init_rank(box[0], size, rank)


def label_sharded(label_fn, cloud: np.ndarray | None, table: torch.Tensor | None, device, src: int = 0):
    """Label an n-point host cloud held by `src` across all ranks.

    label_fn(points_slice (k,>=3) ndarray, table (M,9) tensor on `device`) -> (k,7) float64 ndarray.
    Returns (records of this rank's slice, (lo, hi)).  The cloud is scattered slice by slice, the table is
    broadcast once, and no collective touches the results.
    """
    rank, size = world()
    table = broadcast_table(table, device, src)
    if size == 1:
        return label_fn(cloud, table), (0, len(cloud))
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        meta[0], meta[1] = cloud.shape[0], cloud.shape[1]
    dist.broadcast(meta, src=src)
    n, cols = int(meta[0]), int(meta[1])
    lo, hi = shard_bounds(n, size, rank)
    width = -(-n // size)                       # scatter wants equal shapes: pad ragged slices by one row
    mine = torch.empty((width, cols), dtype=torch.float64, device=device)
    if rank == src:
        full = torch.as_tensor(np.ascontiguousarray(cloud, dtype=np.float64), device=device)
        parts = []
        for r in range(size):
            a, b = shard_bounds(n, size, r)
            part = torch.zeros((width, cols), dtype=torch.float64, device=device)
            part[: b - a] = full[a:b]
            parts.append(part)
        dist.scatter(mine, parts, src=src)
    else:
        dist.scatter(mine, None, src=src)
    return label_fn(mine[: hi - lo].cpu().numpy(), table), (lo, hi)


def noise_cloud_sharded(rows_fn, records: np.ndarray | None, first_point: np.ndarray | None, seed: int, device, src: int = 0):
    """Noisy surface cloud of a QSM (PreProcessing/NoiseDataGeneration.py) across all ranks: `src` holds the per-cylinder
    plan (``records`` (M,14) float64, ``first_point`` (M+1,) int64, see ``tm_noise_cloud``) and broadcasts it once together
    with the seed; every rank then generates its own contiguous rows of the cloud.  The variates are a function of the
    point number, so the union of the slices is the cloud a single process would have produced.

    rows_fn(records tensor, first_point tensor, lo, hi, seed) -> the rows [lo, hi) of the cloud.
    Returns (rows of this rank, (lo, hi)); there is no collective on the data path.
    """
    rank, size = world()
    if size == 1:
        rec = torch.as_tensor(records, dtype=torch.float64, device=device)
        first = torch.as_tensor(first_point, dtype=torch.int64, device=device)
        n = int(first[-1]) if len(first) else 0
        return rows_fn(rec, first, 0, n, seed), (0, n)
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        meta[0], meta[1] = records.shape[0], np.int64(np.uint64(seed & (2 ** 64 - 1)).astype(np.int64))
    dist.broadcast(meta, src=src)
    m = int(meta[0])
    seed = int(meta[1]) & (2 ** 64 - 1)
    if rank == src:
        rec = torch.as_tensor(np.ascontiguousarray(records, dtype=np.float64), device=device)
        first = torch.as_tensor(np.ascontiguousarray(first_point, dtype=np.int64), device=device)
    else:
        rec = torch.empty((m, 14), dtype=torch.float64, device=device)
        first = torch.empty(m + 1, dtype=torch.int64, device=device)
    dist.broadcast(rec, src=src)
    dist.broadcast(first, src=src)
    n = int(first[-1])
    lo, hi = shard_bounds(n, size, rank)
    return rows_fn(rec, first, lo, hi, seed), (lo, hi)


def gather_records(records: np.ndarray, device, dst: int = 0) -> np.ndarray | None:
    """Optional: concatenate every rank's (k,7) records on `dst` in rank order."""
    rank, size = world()
    if size == 1:
        return records
    mine = torch.as_tensor(records, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(size)]
    dist.all_gather(counts, torch.tensor([mine.shape[0]], dtype=torch.int64, device=device))
    width = max(int(c.item()) for c in counts)
    padded = torch.zeros((width, 7), dtype=torch.float64, device=device)
    padded[: mine.shape[0]] = mine
    if rank == dst:
        outs = [torch.empty((width, 7), dtype=torch.float64, device=device) for _ in counts]
        dist.gather(padded, outs, dst=dst)
        return torch.cat([o[: int(c.item())] for o, c in zip(outs, counts)]).cpu().numpy()
    dist.gather(padded, None, dst=dst)
    return None
