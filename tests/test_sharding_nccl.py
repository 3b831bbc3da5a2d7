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
