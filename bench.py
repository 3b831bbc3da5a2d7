#!/usr/bin/env python
"""Benchmark of the nearest-cylinder label + offset path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

Workload (config.workload): BASELINE.json configs[2] — a noise-augmented plot, 10M NoiseDataGeneration-style
points against a 50k-cylinder QSM (10 synthetic trees), variant A (label generation).  One "step" labels the
cloud once: device-resident fp32 points in, device-resident (index, id, distance, offset) out, including
everything that depends on the points (voxel binning / sort, tile kernel, ring / tree search, winner epilogue in input
order); the per-table voxel index and BVH are built once before the timed region (reported as setup_ms).

With N GPUs (default --scaling strong, north_star's split) the ONE 10M-point plot is sharded: rank 0 broadcasts the
cylinder table once over NCCL, rank r labels the contiguous rows shard_bounds(10M, N, r), there is no data-path
collective, `value` = 10M points / the slowest rank's step time.  The same run also reports the weak-scaling figure
(every rank labels a full copy, key "weak") and checks on the hardware that the sharded rows, gathered over NCCL, are
bit-identical to the single-GPU labelling (key "sharded_parity").  --scaling weak gives every rank its own 10M points.

`e2e` goes through the call a user of the reference makes: LabelGenerationCuda.generate_offset_cloud_cuda_batched(
float64 pageable cloud, DataFrame, device) -> (N,7) float64 records, table install and every copy inside the timed
region; `e2e_pinned` is Engine.label_cloud_host on page-locked buffers with the table resident.  At N = 1 the line also
carries `cpu_baseline` (the reference algorithm on the host cores, oracle port) and `reference_cuda` (the reference's ATen
call chain on this GPU, bounded sample, results compared bit for bit).  Prints ONE JSON line on rank 0.

Timing: CUDA events on the launching stream around every step, L2 flushed between steps (256 MiB write),
barrier + synchronize on both sides of the timed loop, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 10_000_000
N_CYLINDERS = 50_000
METRIC = "points/sec nearest-cylinder label+offset"
UNIT = "points/s"
OPS_PER_PAIR = 81                 # fp32 lane-ops per evaluated pair in reference order
BYTES_PER_POINT = 36              # compulsory HBM traffic per point: 12 B xyz in + 24 B index/id/dist/offset out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region of this path is tens of
    milliseconds, shorter than nvidia-smi's sampling period, so NVML is polled in-process from a thread every
    millisecond (same counters as the B200_PROFILING.md nvidia-smi line); nvidia-smi is the fallback."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.thread, self.stop_flag, self.nvml, self.handle = gpu_index, [], None, False, None, None
        self.smax = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((time.time(), float(clk), int(rs)))
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self, t0: float, t1: float) -> dict:
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.nvml is None:
            return self._smi_once()
        inside = [(c, r) for ts, c, r in self.rows if t0 <= ts <= t1]
        if len(inside) < 3:                   # region shorter than a few polls: add the nearest samples around it
            inside = [(c, r) for _, c, r in sorted(self.rows, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:5]]
        reasons = sorted(name for name, bit in self.REASONS.items() if any(r & bit for _, r in inside))
        return {"sm_mhz": float(np.median([c for c, _ in inside])) if inside else None, "sm_max_mhz": self.smax,
                "reasons": reasons, "samples": len(inside), "source": "NVML polled every 1 ms during the timed region"}

    def _smi_once(self) -> dict:
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().split(",")]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                    "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi, one sample right after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}


def build_workload(seed_points: int):
    from treemorph_b200 import synth
    qsm = synth.random_qsm(N_CYLINDERS, seed=1)
    pts = synth.sample_points(qsm, N_POINTS, seed=seed_points)
    return qsm, pts


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host cores (oracle port; the reference is Python/torch and
# does not travel to the GPU box)
# ------------------------------------------------------------------------------------------------------------

def cpu_sample_size(target_s: float, qsm, pts, variant) -> tuple[int, float]:
    """Calibrate on 2k points, then size the sample for ~target_s seconds of CPU work."""
    from oracle import oracle
    from treemorph_b200 import synth
    arrs = synth.cylinder_arrays(qsm, variant.axis_eps)
    t0 = time.perf_counter()
    oracle.label(pts[:2000], *_oracle_args(arrs), variant, threads=host_threads())
    dt = time.perf_counter() - t0
    n = int(min(len(pts), max(2000, 2000 * target_s / max(dt, 1e-4))))
    return n, dt


def _oracle_args(arrs):
    start, radius, length, unit, ids = arrs
    return start, radius, length, unit, ids


def host_threads() -> int:
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask the OS instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_cpu(qsm, pts, n_sample: int, threads: int = 0):
    from oracle import oracle
    from treemorph_b200 import synth
    arrs = synth.cylinder_arrays(qsm)
    threads = threads or host_threads()
    t0 = time.perf_counter()
    res = oracle.label(pts[:n_sample], *_oracle_args(arrs), oracle.VARIANT_A, threads=threads)
    return time.perf_counter() - t0, res


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    qsm, pts = build_workload(2)
    n_s, _ = cpu_sample_size(4.0, qsm, pts, oracle.VARIANT_A)
    for _ in range(args.warmup):
        run_cpu(qsm, pts, min(n_s, 4000))
    total = 0.0
    for _ in range(args.steps):
        dt, _ = run_cpu(qsm, pts, n_s)
        total += dt
    value = n_s * args.steps / total
    cores = host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{N_POINTS} points x {N_CYLINDERS} cylinders, variant A (label generation)",
                   "sample": f"first {n_s} points of the cloud against all {N_CYLINDERS} cylinders per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_s} points x {N_CYLINDERS} cylinders per step, {args.steps} steps, OpenMP {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0



def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each per-call kernel at the default workload, from the
    newest committed ncu capture (profiles/*_dram_traffic.json, written by profiles/summarise.py together with the
    git hash of the build that was profiled).  None when there is no capture."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_dram_traffic.json")))
    if not files:
        return None
    with open(files[-1]) as f:
        d = json.load(f)
    d["source"] = f"profiles/{os.path.basename(files[-1])} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, build {d.get('git', '?')})"
    return d


def stats_lane_ops(stats: dict) -> dict:
    """FP32 lane-operations the counted work stands for: full evaluations in reference order (81) plus, when
    the library reports them, the approximate-distance bounds of the tile kernel at their own documented count."""
    pairs = stats["pairs_evaluated"]
    bounds = stats.get("bound_tests", 0)
    per_bound = stats.get("lane_ops_per_bound", 0)
    return {"total": pairs * OPS_PER_PAIR + bounds * per_bound,
            "detail": {"bound_tests": bounds, "lane_ops_per_bound": per_bound}}

# ------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------

def main():
    global N_POINTS, N_CYLINDERS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="grid", choices=["grid", "brute", "auto"])
    ap.add_argument("--cell", type=float, default=0.0)
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--cylinders", type=int, default=N_CYLINDERS)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: ONE plot of --points points sharded over the ranks (north_star); weak: --points per rank")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-brute", action="store_true", help="skip the exhaustive-kernel FP32 yard-stick")
    ap.add_argument("--skip-e2e", action="store_true", help="skip the host-API end-to-end leg (profiling runs)")
    args = ap.parse_args()
    N_POINTS, N_CYLINDERS = args.points, args.cylinders
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; this path has no CPU fallback"}))
        return 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the one JSON line: NCCL prints its banner ("NCCL version ...", with NCCL_DEBUG=VERSION always on
        # stdout) while the communicator is created, so file descriptor 1 points at stderr for that long
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from treemorph_b200 import api, sharding, synth
    from treemorph_b200.PreProcessing import LabelGenerationCuda as dropin_A
    eng = api.get_engine(dev)              # the engine the drop-in modules use
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    strong = args.scaling == "strong"

    # ---- inputs: rank 0 owns the QSM and broadcasts the packed table once.  strong: ONE cloud, rank r owns the
    #      contiguous rows shard_bounds(N, world, r) of it (LabelGenerationCuda.py:126-133 is the loop being sharded);
    #      weak: every rank samples its own N-point cloud
    t_gen = time.time()
    qsm = synth.random_qsm(N_CYLINDERS, seed=1)
    table = None
    if rank == 0:
        start, radius, length, unit, ids = synth.cylinder_arrays(qsm)
        table = sharding.pack_table(torch.tensor(start), torch.tensor(radius), torch.tensor(length), torch.tensor(unit),
                                    torch.tensor(ids)).to(dev)
    # The table travels through the library's own communicator (C ABI: tm_comm_init_rank + tm_broadcast_cylinders, NCCL over
    # NVLink; torch.distributed only ships the 128-byte NCCL id) and is installed on arrival.  The torch copy of it is kept
    # for the re-installs further down.
    comm_note = "single GPU: tm_broadcast_cylinders is tm_set_cylinders"
    if world > 1:
        try:
            sharding.init_engine_comm(eng)
            comm_note = "tm_comm_init_rank + tm_broadcast_cylinders (NCCL bound by the library), install included"
        except Exception as exc:      # e.g. no loadable libnccl: fall back to torch's process group, and say so
            comm_note = f"torch.distributed broadcast (C-ABI communicator unavailable: {exc})"
        dist.barrier()                      # communicator set-up is not part of the broadcast being timed
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bcast_ms = first_bcast_ms = None
    for attempt in range(2):                # the first collective of a new communicator also sets up its channels
        t_b = time.perf_counter()
        if eng.comm_info()[1] == world:
            if rank == 0:
                s0, r0, l0, u0, i0 = sharding.unpack_table(table)
                eng.broadcast_cylinders(s0, r0, l0, u0, i0, root=0)
            else:
                eng.broadcast_cylinders(root=0)
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t_b) * 1e3
        if attempt == 0:
            first_bcast_ms = bcast_ms
    table = sharding.broadcast_table(table, dev)
    s_t, r_t, l_t, u_t, i_t = sharding.unpack_table(table)
    cloud_host = synth.sample_points(qsm, N_POINTS, seed=2 if strong else 2 + rank)
    lo, hi = sharding.shard_bounds(N_POINTS, world, rank) if strong else (0, N_POINTS)
    n_mine = hi - lo
    pts_host = cloud_host[lo:hi]
    total_points = N_POINTS if strong else world * N_POINTS
    pinned_in = torch.empty((n_mine, 3), dtype=torch.float32, pin_memory=True)
    pinned_in.numpy()[:] = pts_host
    dpts = pinned_in.to(dev, non_blocking=True)

    def new_out(n):
        return {"index": torch.empty(n, dtype=torch.int32, device=dev), "id": torch.empty(n, dtype=torch.int32, device=dev),
                "dist": torch.empty(n, dtype=torch.float32, device=dev), "offset": torch.empty((n, 3), dtype=torch.float32, device=dev)}
    out = new_out(n_mine)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gen_s = time.time() - t_gen

    # ---- table install (per QSM, not per step): pack + solid AABBs + voxel index
    e0.record()
    eng.set_cylinders(s_t, r_t, l_t, u_t, i_t)
    eng.label(dpts[:4096], api.VARIANT_A, mode=args.mode, cell_size=args.cell, want=("id",))     # builds the voxel index
    e1.record()
    torch.cuda.synchronize()
    setup_ms = e0.elapsed_time(e1)

    def timed(points, result, n_steps, sample_clocks=False):
        """warm-up, then n_steps steps: CUDA events on the launching stream, L2 flushed between steps, barrier + synchronize
        on both sides, max over ranks.  Returns (total ms, clocks)."""
        def one():
            eng.label(points, api.VARIANT_A, mode=args.mode, cell_size=args.cell, out=result, want=("index", "id", "dist", "offset"))
        for _ in range(warmup):
            flush.fill_(1)
            one()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        if sample_clocks and rank == 0:
            sampler.start()
            time.sleep(0.05)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        for a, b in evs:
            flush.fill_(1)                      # evict the previous step's working set from the 126 MB L2
            a.record()
            one()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t1 = time.time()
        clocks = sampler.stop(t0, t1) if (sample_clocks and rank == 0) else None
        tm = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item()), clocks

    # ---- timed region
    total_ms, clocks = timed(dpts, out, steps, sample_clocks=True)
    ms_per_step = total_ms / steps
    value = total_points * steps / (total_ms * 1e-3)
    stats = eng.stats()

    # ---- per-phase device time (CUDA events inside the library, same stream)
    eng.set_profiling(True)
    phase_acc = {}
    for _ in range(3):
        flush.fill_(1)
        eng.label(dpts, api.VARIANT_A, mode=args.mode, cell_size=args.cell, out=out, want=("index", "id", "dist", "offset"))
        for k, v in eng.phase_ms().items():
            phase_acc.setdefault(k, []).append(v)
    eng.set_profiling(False)
    phases = {k: float(np.mean(v)) for k, v in phase_acc.items()}
    stats = eng.stats()

    # ---- N > 1, strong: (i) the weak-scaling figure as an extra key (every rank labels the WHOLE cloud), which also gives
    #      every rank the single-GPU answer; (ii) parity on hardware: the sharded rows, gathered over NCCL, against the
    #      single-GPU labelling of the same cloud, bit for bit
    weak = parity = None
    if world > 1 and strong:
        dfull = torch.as_tensor(cloud_host).to(dev)
        out_full = new_out(N_POINTS)
        wms, _ = timed(dfull, out_full, steps)
        weak = {"value": world * N_POINTS * steps / (wms * 1e-3), "ms_per_step": wms / steps, "unit": UNIT,
                "note": "every rank labels a full copy of the cloud (no sharding)"}
        eng.label(dpts, api.VARIANT_A, mode=args.mode, cell_size=args.cell, out=out, want=("index", "id", "dist", "offset"))
        width = -(-N_POINTS // world)
        same = {}
        for k in ("index", "id", "dist", "offset"):
            mine = out[k].reshape(n_mine, -1)
            pad = torch.zeros((width, mine.shape[1]), dtype=mine.dtype, device=dev)
            pad[:n_mine] = mine
            parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, parts, dst=0)
            if rank == 0:
                rows = torch.cat([p[: sharding.shard_bounds(N_POINTS, world, r)[1] - sharding.shard_bounds(N_POINTS, world, r)[0]]
                                  for r, p in enumerate(parts)])
                ref = out_full[k].reshape(N_POINTS, -1)
                same[k] = bool(torch.equal(rows.view(torch.int32), ref.view(torch.int32)))      # bit patterns (NaN-safe)
        if rank == 0:
            parity = {"sharded_vs_single_gpu_bitwise": same, "rows": N_POINTS, "collective": "NCCL gather of every rank's rows to rank 0"}
        del dfull, out_full
        eng.label(dpts, api.VARIANT_A, mode=args.mode, cell_size=args.cell, out=out, want=("index", "id", "dist", "offset"))
        stats = eng.stats()

    # ---- end to end (i): the call a user of the reference makes — generate_offset_cloud_cuda_batched(float64 pageable
    #      cloud, DataFrame, device) of the LabelGenerationCuda drop-in (reference :113), table install included
    e2e_steps = 0 if args.skip_e2e else max(2, min(steps, 5))
    cloud64 = np.ascontiguousarray(pts_host, dtype=np.float64)            # pageable, the dtype np.load of a cloud file yields
    qsm_df = synth.qsm_dataframe(qsm)
    e2e_value = e2e_ok = None
    rec = None
    if e2e_steps:
        for _ in range(2):
            rec = dropin_A.generate_offset_cloud_cuda_batched(cloud64, qsm_df, dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            rec = dropin_A.generate_offset_cloud_cuda_batched(cloud64, qsm_df, dev)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = total_points * e2e_steps / float(te.item())
        e2e_ok = bool((rec[:, 6] == out["id"].cpu().numpy()).all() and
                      np.array_equal(rec[:, 3:6].astype(np.float32), out["offset"].cpu().numpy(), equal_nan=True)
                      and np.array_equal(rec[:, :3], cloud64))
    pipe = eng.host_pipeline_info() if e2e_steps else {"d2h_bytes_per_point": 16, "host_threads": None}
    del rec

    # ---- end to end (ii): Engine.label_cloud_host on page-locked fp32 / page-locked records (table already installed)
    eng.set_cylinders(s_t, r_t, l_t, u_t, i_t)
    rec_pinned = torch.empty((n_mine, 7), dtype=torch.float64, pin_memory=True)
    rec_np, cloud_np = rec_pinned.numpy(), pinned_in.numpy()
    pinned_value = None
    if e2e_steps:
        for _ in range(2):
            eng.label_cloud_host(cloud_np, api.VARIANT_A, mode=args.mode, cell_size=args.cell, out=rec_np)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.label_cloud_host(cloud_np, api.VARIANT_A, mode=args.mode, cell_size=args.cell, out=rec_np)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        pinned_value = total_points * e2e_steps / float(te.item())
    pipe_pinned = eng.host_pipeline_info() if e2e_steps else pipe
    # host memory system: the end-to-end call is bound by it (profiles/r01i_host_pipeline.md).  Bytes the host cores and the
    # DMA engines move per point of a float64 cloud: read 24 (staging) + write 12 (page-locked float32) + DMA read 12 +
    # DMA write 16 ({offset, id}) + read 16 + read 24 (xyz for the record) + write 56 (the (N,7) float64 record) = 160
    host_bw = eng.host_bandwidth() if e2e_steps else None
    host_view = None
    if e2e_steps and e2e_value:
        per_point = 160
        host_view = {"host_bytes_per_point": per_point, "achieved_GBps": per_point * e2e_value / 1e9,
                     "copy_peak_GBps": host_bw["bytes_per_s"] / 1e9, "threads": host_bw["threads"],
                     "note": "STREAM-style copy (read + write) with the library's host workers; ranks of one node share the host"}
        host_view["frac"] = host_view["achieved_GBps"] / host_view["copy_peak_GBps"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines
    hbm_peak, peak_kind = load_peaks()
    fp32_peak = eng.fp32_peak()
    dom = max(("evaluate", "tree", "exhaustive", "pending", "bin", "scatter", "scan", "epilogue"), key=lambda k: phases.get(k, 0.0))
    dom_ms = phases.get(dom, 0.0) or ms_per_step
    achieved_gbs = BYTES_PER_POINT * n_mine / (dom_ms * 1e-3) / 1e9
    pairs = stats["pairs_evaluated"]
    default_workload = (N_POINTS, N_CYLINDERS) == (10_000_000, 50_000) and args.mode == "grid" and n_mine == N_POINTS
    traffic = load_ncu_traffic() if default_workload else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak,
                "traffic": (traffic or {}).get("kernels", {}).get(dom),
                "traffic_source": (traffic or {}).get("source"), "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "kernel_ms": dom_ms,
                "note": "dominant kernel is FP32-issue bound, not HBM bound: see fp32_roofline"}
    lane_ops = stats_lane_ops(stats)
    fp32_roofline = {"kernel": dom, "pairs_evaluated": pairs, "lane_ops_per_pair": OPS_PER_PAIR, **lane_ops["detail"],
                     "achieved_lane_ops_per_s": lane_ops["total"] / (dom_ms * 1e-3), "peak_lane_ops_per_s": fp32_peak,
                     "frac": lane_ops["total"] / (dom_ms * 1e-3) / fp32_peak if fp32_peak else None,
                     "peak_source": "FFMA/FADD+FMUL chain probe in this run"}

    # whole step against HBM (compulsory bytes per point over the step), and the brute-force-equivalent pair
    # rate N*M/t -- a speed-up figure of the pruning, not a roofline fraction
    step_gbs = BYTES_PER_POINT * n_mine / (ms_per_step * 1e-3) / 1e9
    step_view = {"compulsory_bytes_per_point": BYTES_PER_POINT, "achieved_GBps": step_gbs, "hbm_frac": step_gbs / hbm_peak,
                 "dram_bytes_all_kernels": sum((traffic or {}).get("kernels", {}).values()) or None,
                 "brute_force_equivalent_pairs_per_s": float(total_points) * N_CYLINDERS / (ms_per_step * 1e-3),
                 "brute_force_equivalent_lane_ops_per_s": float(total_points) * N_CYLINDERS * OPS_PER_PAIR / (ms_per_step * 1e-3),
                 "pairs_evaluated_per_point": pairs / n_mine, "cull_tests_per_point": stats["cull_tests"] / n_mine}

    # ---- exhaustive kernel as the FP32 yard-stick (pairs = N*M exactly)
    brute = None
    if not args.skip_brute:
        nb = min(n_mine, 400_000)
        eng.label(dpts[:nb], api.VARIANT_A, mode="brute", want=("id",))
        torch.cuda.synchronize()
        e0.record()
        eng.label(dpts[:nb], api.VARIANT_A, mode="brute", want=("index", "id", "dist", "offset"))
        e1.record()
        torch.cuda.synchronize()
        bms = e0.elapsed_time(e1)
        bp = nb * N_CYLINDERS
        brute = {"points": nb, "pairs": bp, "ms": bms, "pairs_per_s": bp / (bms * 1e-3),
                 "fp32_frac": bp * OPS_PER_PAIR / (bms * 1e-3) / fp32_peak if fp32_peak else None}

    # ---- reference algorithm on the host cores, bounded sample, same run
    cpu = reference_cuda = None
    if not args.skip_cpu and world == 1:          # the contract: rank 0 at N = 1 only
        from oracle import oracle
        oracle.build()
        n_s, _ = cpu_sample_size(12.0, qsm, pts_host, oracle.VARIANT_A)
        dt, res = run_cpu(qsm, pts_host, n_s)
        same = bool((res["id"] == out["id"][:n_s].cpu().numpy()).all())
        cpu = {"value": n_s / dt, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": f"first {n_s} points x {N_CYLINDERS} cylinders, {dt:.1f} s, OpenMP {host_threads()} threads "
                         f"of {os.cpu_count()} host cpus; ids equal to the GPU result: {same}"}
        # the reference's own deployment is device=cuda (Modules/Utils.py:146-158): its ATen call chain, restated op for op
        # (oracle/torch_mirror.py, checker code like the CPU port), on THIS GPU with the DataFrame table layout and the
        # reference's batch size, on a bounded sample of the same cloud — the GPU-versus-GPU comparator
        try:
            from oracle import torch_mirror
            n_r = min(len(pts_host), 20_480)
            torch_mirror.label(pts_host[:2048], qsm, dev, 1e-6, 0.0, 0.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref = torch_mirror.label(pts_host[:n_r], qsm, dev, 1e-6, 0.0, 0.0)
            torch.cuda.synchronize()
            dt_r = time.perf_counter() - t0
            reference_cuda = {"value": n_r / dt_r, "unit": UNIT, "kind": "torch mirror of the reference's ATen chain on this GPU",
                              "sample": f"first {n_r} points x {N_CYLINDERS} cylinders, batch_size 1024, {dt_r:.2f} s",
                              "ids_equal_to_ours": bool((ref["id"] == out["id"][:n_r].cpu().numpy()).all()),
                              "dist_bitwise_equal_to_ours": bool(np.array_equal(ref["dist"], out["dist"][:n_r].cpu().numpy(), equal_nan=True)),
                              "speedup_device_resident": value / (n_r / dt_r), "speedup_e2e": (e2e_value or 0.0) / (n_r / dt_r)}
        except Exception as exc:           # e.g. out of memory on a smaller part: the comparator is optional
            reference_cuda = {"unavailable": str(exc)[:200]}

    launches_per_step = stats.get("launches") or {"grid": 11, "auto": 11, "brute": 2}[args.mode]
    shard_txt = (f"ONE plot of {N_POINTS} points sharded over {world} GPU(s) ({n_mine} rows on rank 0)" if strong
                 else f"{N_POINTS} points per GPU")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{shard_txt} x {N_CYLINDERS} cylinders, variant A (label generation), "
                               f"random QSM plot + NoiseDataGeneration-style cloud", "mode": args.mode,
                   "cell_size_m": stats.get("cell_size"), "l2": "flushed between steps (256 MiB write)",
                   "parallelism": f"points sharded x{world}, cylinder table broadcast once over NCCL, no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_mine * 12, "d2h_bytes_per_step": n_mine * pipe["d2h_bytes_per_point"],
                "api": "PreProcessing.LabelGenerationCuda.generate_offset_cloud_cuda_batched(float64 pageable cloud, DataFrame, device) "
                       "-> (N,7) float64 records; the DataFrame is read and compared with the installed table inside every call (same values: no re-install); "
                       "host workers round the cloud to float32 into page-locked staging, 12 B/point cross PCIe",
                "host_assembly_threads": pipe["host_threads"], "bytes_are": "per rank", "host_memory": host_view,
                "steps": e2e_steps, "checked": e2e_ok},
        "e2e_pinned": {"value": pinned_value, "unit": UNIT, "h2d_bytes_per_step": n_mine * 12,
                       "d2h_bytes_per_step": n_mine * pipe_pinned["d2h_bytes_per_point"],
                       "api": "Engine.label_cloud_host (tm_label_cloud_host): page-locked fp32 cloud -> page-locked (N,7) float64 records, table resident",
                       "host_assembly_threads": pipe_pinned["host_threads"]},
        "gpu_launches": launches_per_step * steps,
        "roofline": roofline, "fp32_roofline": fp32_roofline, "step_view": step_view, "brute_force_yardstick": brute,
        "cpu_baseline": cpu, "reference_cuda": reference_cuda, "weak": weak, "sharded_parity": parity,
        "phases_ms": phases, "stats": stats, "setup_ms": setup_ms, "table_broadcast_ms": bcast_ms, "table_broadcast_first_ms": first_bcast_ms, "table_broadcast": comm_note,
        "input_generation_s": gen_s,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
