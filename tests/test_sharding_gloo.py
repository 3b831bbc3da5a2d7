"""world_size-2 test of the multi-GPU plumbing on CPU (gloo): one table broadcast, contiguous point slices,
no collective on the results.  The per-rank labeller here is the CPU oracle (tests may use it as the checker);
on GPUs it is the CUDA engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from treemorph_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_table_roundtrip_keeps_id_bits():
    ids = np.array([0, 5, -7, 2**31 - 1, 123456789], dtype=np.int32)
    m = len(ids)
    rng = np.random.default_rng(0)
    t = sharding.pack_table(rng.random((m, 3)), rng.random(m), rng.random((m, 1)), rng.random((m, 3)), ids)
    s, r, l, u, i = sharding.unpack_table(t)
    assert t.shape == (m, 9) and l.shape == (m, 1)
    assert (i.numpy() == ids).all()


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from treemorph_b200 import synth
        dev = torch.device("cpu")
        cloud = table = None
        q = synth.random_qsm(150, seed=3, id_offset=10)
        if rank == 0:                                   # only the source rank holds the inputs
            cloud = synth.sample_points(q, 1001, seed=4).astype(np.float64)
            start, radius, length, unit, ids = synth.cylinder_arrays(q)
            table = sharding.pack_table(start, radius, length, unit, ids)

        def label_fn(points, tab):
            s, r, l, u, i = sharding.unpack_table(tab)
            res = oracle.label(points, s.numpy(), r.numpy(), l.numpy(), u.numpy(), i.numpy(), oracle.VARIANT_A)
            out = np.zeros((len(points), 7))
            out[:, :3], out[:, 3:6], out[:, 6] = points[:, :3], res["offset"], res["id"]
            return out

        rec, (lo, hi) = sharding.label_sharded(label_fn, cloud, table, dev)
        assert (lo, hi) == sharding.shard_bounds(1001, world, rank) and rec.shape == (hi - lo, 7)
        full = sharding.gather_records(rec, dev)
        if rank == 0:
            want = oracle.label_cloud(cloud, q, oracle.VARIANT_A)
            assert (full == want).all()
            np.save(os.path.join(tmpdir, "ok.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_labelling_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")


def _noise_worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import noise_cloud as nc
        from treemorph_b200 import synth
        from treemorph_b200.PreProcessing import NoiseDataGeneration as N
        seed = 0xFEDC_BA98_7654_3210                          # needs all 64 bits to survive the broadcast
        df = synth.qsm_dataframe(synth.random_qsm(80, seed=6))
        plan = N.cylinder_plan(df)
        records = first = None
        if rank == 0:                                          # only the source rank holds the plan
            records, first = plan.records, plan.first_point

        def rows_fn(rec, first_point, lo, hi, sd):            # CPU stand-in for Engine.noise_cloud(point0=lo, n=hi-lo)
            rec = rec.numpy()
            p = nc.Plan(rec[:, 0:3], rec[:, 13], rec[:, 12], rec[:, 3:12].reshape(-1, 3, 3), np.diff(first_point.numpy()))
            theta, z, noise = nc.philox_variates(p, sd, first=lo, n=hi - lo)
            cid = nc.owners(p)[lo:hi]
            rho = p.radius[cid] + noise
            local = np.stack([rho * np.cos(theta), rho * np.sin(theta), z], axis=1)
            return np.einsum("nij,nj->ni", p.rot[cid], local) + p.start[cid]

        rows, (lo, hi) = sharding.noise_cloud_sharded(rows_fn, records, first, seed, torch.device("cpu"))
        assert (lo, hi) == sharding.shard_bounds(plan.n_points, world, rank) and rows.shape == (hi - lo, 3)
        op = nc.plan(df[["startX", "startY", "startZ"]].values, df[["endX", "endY", "endZ"]].values, df["radius"].values)
        want = nc.place(op, *nc.philox_variates(op, seed))
        assert np.array_equal(rows, want[lo:hi])               # the slices tile the single-process cloud exactly
        np.save(os.path.join(tmpdir, f"ok{rank}.npy"), np.array([hi - lo]))
    finally:
        dist.destroy_process_group()


def test_two_rank_noise_cloud_tiles_the_single_process_cloud(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_noise_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")
