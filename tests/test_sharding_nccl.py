"""The sharded product path on hardware: two ranks, one GPU each, NCCL.  `sharding.label_sharded` (table broadcast +
point scatter from rank 0, each rank labels its rows with the CUDA engine) followed by `sharding.gather_records` must
reproduce the single-GPU (N,7) records bit for bit (the loop being sharded: LabelGenerationCuda.py:126-133).

Needs two visible GPUs (`gpurun --gpus 2`); on a one-GPU box it is skipped — bench.py repeats the same check inside every
N > 1 run ("sharded_parity" in its JSON line), which is what the driver's scaling runs exercise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from treemorph_b200 import api, sharding, synth
        eng = api.Engine(dev)
        n, m = 300_001, 5000                       # ragged split: the ranks get 150 001 and 150 000 rows
        q = synth.random_qsm(m, seed=31, id_offset=1000)
        cloud = table = None
        if rank == 0:                              # only the source rank holds the inputs
            cloud = synth.sample_points(q, n, seed=32).astype(np.float64)
            start, radius, length, unit, ids = synth.cylinder_arrays(q)
            table = sharding.pack_table(start, radius, length, unit, ids).to(dev)

        def label_fn(points, tab):
            s, r, l, u, i = sharding.unpack_table(tab)
            eng.set_cylinders(s, r, l, u, i)
            return eng.label_cloud_host(points, api.VARIANT_A)

        rec, (lo, hi) = sharding.label_sharded(label_fn, cloud, table, dev)
        assert (lo, hi) == sharding.shard_bounds(n, world, rank) and rec.shape == (hi - lo, 7)
        full = sharding.gather_records(np.ascontiguousarray(rec), dev)
        if rank == 0:
            want = label_fn(cloud, table)          # the same cloud on ONE GPU
            assert full.shape == want.shape
            assert np.array_equal(full.view(np.int64), want.view(np.int64)), "sharded records differ from the single-GPU records"
            from oracle import oracle             # and a slice of it against the checker
            ref = oracle.label_cloud(cloud[:4000], q, oracle.VARIANT_A)
            assert np.array_equal(full[:4000], ref, equal_nan=True)
            np.save(os.path.join(tmpdir, "ok.npy"), np.array([1]))
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_sharded_labelling_matches_single_gpu(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")


def _cabi_worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # only ships the 128-byte NCCL id
    try:
        from treemorph_b200 import api, sharding, synth
        from oracle import oracle
        eng = api.Engine(dev)
        sharding.init_engine_comm(eng)
        assert eng.comm_info() == (rank, world)
        q = synth.random_qsm(3000, seed=41, id_offset=7)
        pts = synth.sample_points(q, 50_000, seed=42)
        if rank == 0:                              # only the root holds the table
            start, radius, length, unit, ids = synth.cylinder_arrays(q)
            m = eng.broadcast_cylinders(torch.tensor(start, device=dev), torch.tensor(radius, device=dev), torch.tensor(length, device=dev),
                                        torch.tensor(unit, device=dev), torch.tensor(ids, device=dev), root=0)
        else:
            m = eng.broadcast_cylinders(root=0)
        assert m == 3000
        lo, hi = sharding.shard_bounds(len(pts), world, rank)
        got = eng.label(torch.tensor(pts[lo:hi], device=dev), api.VARIANT_A, mode="grid")
        start, radius, length, unit, ids = synth.cylinder_arrays(q)
        want = oracle.label(pts[lo:hi], start, radius, length, unit, ids, oracle.VARIANT_A)
        assert (got["id"].cpu().numpy() == want["id"]).all() and np.array_equal(got["dist"].cpu().numpy(), want["dist"], equal_nan=True)
        assert np.array_equal(got["offset"].cpu().numpy(), want["offset"], equal_nan=True)
        eng.comm_destroy()
        eng.close()
        np.save(os.path.join(tmpdir, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_c_abi_communicator_broadcasts_and_installs_the_table(tmp_path):
    """tm_comm_unique_id / tm_comm_init_rank / tm_broadcast_cylinders: rank 1 never sees the QSM, labels its rows against
    the table NCCL delivered, and matches the oracle bit for bit."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_cabi_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")


def test_broadcast_without_a_communicator_is_an_install():
    from treemorph_b200 import api, synth
    from oracle import oracle
    dev = torch.device("cuda", 0)
    eng = api.Engine(dev)
    q = synth.random_qsm(500, seed=43)
    start, radius, length, unit, ids = synth.cylinder_arrays(q)
    assert eng.comm_info() == (0, 1)
    assert eng.broadcast_cylinders(torch.tensor(start, device=dev), torch.tensor(radius, device=dev), torch.tensor(length, device=dev),
                                   torch.tensor(unit, device=dev), torch.tensor(ids, device=dev)) == 500
    pts = synth.sample_points(q, 5000, seed=44)
    got = eng.label(torch.tensor(pts, device=dev), api.VARIANT_A)
    want = oracle.label(pts, start, radius, length, unit, ids, oracle.VARIANT_A)
    assert (got["id"].cpu().numpy() == want["id"]).all()
    eng.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_one_process_two_devices_init_all_and_broadcast_all():
    """tm_comm_init_all / tm_broadcast_cylinders_all: one process, one handle per device; the table lives on device 0 only and
    both devices label their rows of the cloud bit-identically to the oracle."""
    from treemorph_b200 import api, sharding, synth
    from oracle import oracle
    engines = [api.Engine(torch.device("cuda", i)) for i in range(2)]
    api.comm_init_all(engines)
    assert [e.comm_info() for e in engines] == [(0, 2), (1, 2)]
    q = synth.random_qsm(2500, seed=51, id_offset=3)
    start, radius, length, unit, ids = synth.cylinder_arrays(q)
    d0 = engines[0].device
    api.broadcast_cylinders_all(engines, torch.tensor(start, device=d0), torch.tensor(radius, device=d0), torch.tensor(length, device=d0),
                                torch.tensor(unit, device=d0), torch.tensor(ids, device=d0), root_index=0)
    pts = synth.sample_points(q, 40_000, seed=52)
    want = oracle.label(pts, start, radius, length, unit, ids, oracle.VARIANT_A)
    for r, e in enumerate(engines):
        lo, hi = sharding.shard_bounds(len(pts), 2, r)
        with torch.cuda.device(e.device):
            got = e.label(torch.tensor(pts[lo:hi], device=e.device), api.VARIANT_A, mode="grid")
            torch.cuda.synchronize()
        assert (got["id"].cpu().numpy() == want["id"][lo:hi]).all()
        assert np.array_equal(got["dist"].cpu().numpy(), want["dist"][lo:hi], equal_nan=True)
    for e in engines:
        e.comm_destroy()
        e.close()
