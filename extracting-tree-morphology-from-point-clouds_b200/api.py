"""Device-level Python API over the C-ABI library: torch tensors in, torch tensors out.

torch is used for device memory and streams only; every computation happens in
``libtreemorph_nn.so`` (hand-written sm_100a CUDA).  The functions here mirror the reference's
argument conventions for the hot path:

* ``closest_cylinder_cuda_batch(points, start, radius, axis_length, axis_unit, IDs, device)``
  (``PreProcessing/LabelGenerationCuda.py:20``, ``Modules/Projection.py:19``)  → ``Engine.label``
* the cylinder preparation of ``generate_offset_cloud_cuda_batched`` (``LabelGenerationCuda.py:117-123``,
  ``Projection.py:121-132``) → ``Engine.prepare``
* the ``(N,7)`` float64 record (``LabelGenerationCuda.py:114,131-133``) → ``Engine.assemble`` /
  ``Engine.label_cloud_host``
"""
from __future__ import annotations

import ctypes
import os
import weakref
from dataclasses import dataclass

import numpy as np
import torch

from . import binding as B


@dataclass(frozen=True)
class Variant:
    """The two parameterisations of the reference kernel."""
    name: str
    perp_atol: float      # LabelGenerationCuda.py:51 (1e-6)  /  Projection.py:50 (1e-3)
    norm_eps: float       # Projection.py:60-62 (1e-8); 0 = unguarded (LabelGenerationCuda.py:58-60)
    axis_eps: float       # Projection.py:129-131 (1e-8); 0 = unguarded (LabelGenerationCuda.py:123)


VARIANT_A = Variant("A", 1e-6, 0.0, 0.0)       # PreProcessing/LabelGenerationCuda.py
VARIANT_B = Variant("B", 1e-3, 1e-8, 1e-8)     # Modules/Projection.py
VARIANTS = {"A": VARIANT_A, "B": VARIANT_B}
SMALL_TABLE_MAX = 3072      # cylinders per tm_proximity_flags_host call (csrc/tm_core.cuh: SMALL_MAX_M)
MODES = {"auto": B.TM_MODE_AUTO, "brute": B.TM_MODE_BRUTE, "grid": B.TM_MODE_GRID}


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("treemorph_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


_pinned_live = 0            # bytes of page-locked result arrays currently alive (Engine._new_records)


def _pinned_released(nbytes: int) -> None:
    global _pinned_live
    _pinned_live -= nbytes


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else int(t.data_ptr())


_engine_serial = 0          # every Engine gets its own number: caches keyed on "this engine's n-th install" cannot mix engines up


class Engine:
    """One C-ABI handle (one device, one host thread)."""

    def __init__(self, device: torch.device | int | str | None = None):
        _require_cuda()
        self._lib = B.load()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError(f"Engine needs a cuda device, got {dev}")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        h = ctypes.c_void_p()
        B.check(self._lib, None, self._lib.tm_create(dev.index, ctypes.byref(h)))
        self._h = h
        global _engine_serial
        _engine_serial += 1
        self.serial = _engine_serial
        self.m = 0
        self.installs = 0          # bumped by every set_cylinders: callers that cache "my table is installed" compare it
        self.installs_cloud = 0    # bumped by every upload_cloud: callers that cache "my cloud is resident" compare it

    # -- lifetime ----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.tm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status: int) -> None:
        B.check(self._lib, self._h, status)

    def _f32(self, t, cols: int | None = None) -> torch.Tensor:
        t = torch.as_tensor(t)
        if t.device != self.device or t.dtype != torch.float32:
            t = t.to(device=self.device, dtype=torch.float32)
        if cols is not None and (t.dim() != 2 or t.shape[1] != cols):
            raise ValueError(f"expected a (M,{cols}) tensor, got {tuple(t.shape)}")
        return t

    # -- cylinder side -----------------------------------------------------------------------
    def prepare(self, start, end, variant: Variant = VARIANT_A, norm_fma: bool = False):
        """``axis_length`` (M,1) and ``axis_unit`` (M,3) as generate_offset_cloud_cuda_batched builds them."""
        start = self._f32(start, 3)
        end = self._f32(end, 3)
        m = start.shape[0]
        length = torch.empty((m, 1), dtype=torch.float32, device=self.device)
        unit = torch.empty((m, 3), dtype=torch.float32, device=self.device)
        self._check(self._lib.tm_prepare_cylinders(
            self._h, _ptr(start), start.stride(0), start.stride(1), _ptr(end), end.stride(0), end.stride(1), m,
            float(variant.axis_eps), int(bool(norm_fma)), _ptr(length), _ptr(unit), _stream_ptr(self.device)))
        return length, unit

    def set_cylinders(self, start, radius, axis_length, axis_unit, ids=None) -> None:
        """Install the cylinder table (any strides; Fortran-ordered DataFrame tensors are read in place)."""
        start = self._f32(start, 3)
        unit = self._f32(axis_unit, 3)
        m = start.shape[0]
        length = self._f32(axis_length).reshape(-1)
        radius = self._f32(radius).reshape(-1)
        if unit.shape[0] != m or length.numel() != m or radius.numel() != m:
            raise ValueError("cylinder arrays disagree on M")
        if ids is not None:
            ids = torch.as_tensor(ids)
            if ids.device != self.device or ids.dtype != torch.int32:
                ids = ids.to(device=self.device, dtype=torch.int32)
            ids = ids.reshape(-1)
            if ids.numel() != m:
                raise ValueError("IDs disagree on M")
        self._check(self._lib.tm_set_cylinders(
            self._h, _ptr(start), start.stride(0), start.stride(1), _ptr(unit), unit.stride(0), unit.stride(1),
            _ptr(length), length.stride(0) if m else 1, _ptr(radius), radius.stride(0) if m else 1,
            _ptr(ids), (ids.stride(0) if m else 1) if ids is not None else 1, m, _stream_ptr(self.device)))
        self.m = m
        self.installs += 1

    # -- multi-GPU: communicator + table broadcast through the C ABI (NCCL bound by the library) ---------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclGetUniqueId (128 bytes) — rank 0 makes it and ships it to the other ranks."""
        lib = B.load()
        buf = ctypes.create_string_buffer(B.TM_COMM_ID_BYTES)
        B.check(lib, None, lib.tm_comm_unique_id(buf))
        return buf.raw

    def comm_init_rank(self, unique_id: bytes, nranks: int, rank: int) -> None:
        if len(unique_id) != B.TM_COMM_ID_BYTES:
            raise ValueError("unique_id must be the 128 bytes of comm_unique_id()")
        self._check(self._lib.tm_comm_init_rank(self._h, unique_id, int(nranks), int(rank)))

    def comm_destroy(self) -> None:
        self._check(self._lib.tm_comm_destroy(self._h))

    def comm_info(self) -> tuple[int, int]:
        r, n = ctypes.c_int32(), ctypes.c_int32()
        self._check(self._lib.tm_comm_info(self._h, ctypes.byref(r), ctypes.byref(n)))
        return int(r.value), int(n.value)

    def broadcast_cylinders(self, start=None, radius=None, axis_length=None, axis_unit=None, ids=None, root: int = 0) -> int:
        """Replicate the root's cylinder table on every rank of the communicator and install it (tm_broadcast_cylinders).
        The arguments are read on ``root`` only.  Returns M."""
        rank, size = self.comm_info()
        if rank == root or size == 1:
            start = self._f32(start, 3)
            unit = self._f32(axis_unit, 3)
            m = start.shape[0]
            length = self._f32(axis_length).reshape(-1)
            radius = self._f32(radius).reshape(-1)
            if ids is not None:
                ids = torch.as_tensor(ids)
                if ids.device != self.device or ids.dtype != torch.int32:
                    ids = ids.to(device=self.device, dtype=torch.int32)
                ids = ids.reshape(-1)
            args = (_ptr(start), start.stride(0), start.stride(1), _ptr(unit), unit.stride(0), unit.stride(1),
                    _ptr(length), length.stride(0) if m else 1, _ptr(radius), radius.stride(0) if m else 1,
                    _ptr(ids), (ids.stride(0) if m else 1) if ids is not None else 1, m)
        else:
            args = (None, 3, 1, None, 3, 1, None, 1, None, 1, None, 1, 0)
        self._check(self._lib.tm_broadcast_cylinders(self._h, *args, int(root), _stream_ptr(self.device)))
        self.installs += 1
        self.m = int(self.stats_m())
        return self.m

    def stats_m(self) -> int:
        m = ctypes.c_int64()
        self._check(self._lib.tm_cylinder_count(self._h, ctypes.byref(m)))
        return int(m.value)

    # -- point side --------------------------------------------------------------------------
    def _params(self, variant: Variant, move_to_mantle: bool, norm_fma: bool, mode: str, cell_size: float) -> B.TmParams:
        p = B.TmParams()
        p.perp_atol = float(variant.perp_atol)
        p.norm_eps = float(variant.norm_eps)
        p.move_to_mantle = int(bool(move_to_mantle))
        p.norm_fma = int(bool(norm_fma))
        p.mode = MODES[mode]
        p.cell_size = float(cell_size)
        return p

    def label(self, points: torch.Tensor, variant: Variant = VARIANT_A, move_to_mantle: bool = True,
              norm_fma: bool = False, mode: str = "auto", cell_size: float = 0.0,
              want=("index", "id", "dist", "offset"), out: dict | None = None) -> dict:
        """closest_cylinder_cuda_batch for device-resident points → dict of device tensors."""
        if points.device != self.device or points.dtype != torch.float32:
            points = points.to(device=self.device, dtype=torch.float32)
        if points.dim() != 2 or points.shape[1] < 3:
            raise ValueError(f"points must be (N, >=3), got {tuple(points.shape)}")
        if points.shape[0] > 1 and points.stride(1) != 1:
            points = points.contiguous()
        n = points.shape[0]
        row_stride = points.stride(0) if n > 1 else max(3, points.shape[1])
        shapes = {"index": ((n,), torch.int32), "id": ((n,), torch.int32), "dist": ((n,), torch.float32),
                  "offset": ((n, 3), torch.float32), "radius": ((n,), torch.float32)}
        res = {}
        for k in want:
            if out is not None and k in out:
                t = out[k]
                if t.shape != shapes[k][0] or t.dtype != shapes[k][1] or not t.is_contiguous() or t.device != self.device:
                    raise ValueError(f"out[{k!r}] has the wrong shape / dtype / device")
                res[k] = t
            else:
                res[k] = torch.empty(shapes[k][0], dtype=shapes[k][1], device=self.device)
        prm = self._params(variant, move_to_mantle, norm_fma, mode, cell_size)
        self._check(self._lib.tm_label_points(
            self._h, _ptr(points), n, row_stride, ctypes.byref(prm), _ptr(res.get("index")), _ptr(res.get("id")),
            _ptr(res.get("dist")), _ptr(res.get("offset")), _ptr(res.get("radius")), _stream_ptr(self.device)))
        return res

    def assemble(self, cloud: torch.Tensor, offset: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        """(N,7) float64 [xyz, offset, ID] on the device; xyz keeps the cloud's own precision."""
        if cloud.dtype not in (torch.float32, torch.float64):
            cloud = cloud.to(torch.float64)
        if cloud.device != self.device:
            cloud = cloud.to(self.device)
        if cloud.shape[0] > 1 and cloud.stride(1) != 1:
            cloud = cloud.contiguous()
        n = cloud.shape[0]
        out = torch.empty((n, 7), dtype=torch.float64, device=self.device)
        self._check(self._lib.tm_assemble_records(
            self._h, _ptr(cloud), B.TM_F32 if cloud.dtype == torch.float32 else B.TM_F64, n,
            cloud.stride(0) if n > 1 else max(3, cloud.shape[1]), _ptr(offset.contiguous()), _ptr(ids.contiguous()),
            _ptr(out), _stream_ptr(self.device)))
        return out

    def label_cloud_host(self, cloud: np.ndarray, variant: Variant = VARIANT_A, move_to_mantle: bool = True,
                         norm_fma: bool = False, mode: str = "auto", cell_size: float = 0.0,
                         out: np.ndarray | None = None, want_dist: bool = False, tail=None):
        """generate_offset_cloud_cuda_batched for a host cloud: pipelined H2D / label / assemble / D2H.

        ``tail``: values of extra columns appended to every row (the drivers' four dummy feature columns); ``out`` is then
        ``(N, 7 + len(tail))`` and may be a memory-mapped ``.npy`` file."""
        cloud = np.asarray(cloud)
        if cloud.ndim != 2 or cloud.shape[1] < 3:
            raise ValueError(f"cloud must be (N, >=3), got {cloud.shape}")
        if cloud.dtype not in (np.float32, np.float64):
            cloud = cloud.astype(np.float64)
        if cloud.shape[0] > 1 and (cloud.strides[1] != cloud.itemsize or cloud.strides[0] % cloud.itemsize
                                   or cloud.strides[0] < 3 * cloud.itemsize):
            cloud = np.ascontiguousarray(cloud)
        n = cloud.shape[0]
        width = 7 + (len(tail) if tail is not None else 0)
        if out is None:
            out = self._new_records(n, width)
        if out.shape != (n, width) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float64 (N,{width}) array")
        dist = np.empty(n, dtype=np.float32) if want_dist else None
        prm = self._params(variant, move_to_mantle, norm_fma, mode, cell_size)
        row_stride = cloud.strides[0] // cloud.itemsize if n > 1 else max(3, cloud.shape[1])
        if width == 7:
            self._check(self._lib.tm_label_cloud_host(
                self._h, cloud.ctypes.data, B.TM_F32 if cloud.dtype == np.float32 else B.TM_F64, n, row_stride,
                ctypes.byref(prm), out.ctypes.data, dist.ctypes.data if dist is not None else None))
        else:
            tl = (ctypes.c_double * (width - 7))(*[float(v) for v in tail])
            self._check(self._lib.tm_label_cloud_host_wide(
                self._h, cloud.ctypes.data, B.TM_F32 if cloud.dtype == np.float32 else B.TM_F64, n, row_stride,
                ctypes.byref(prm), out.ctypes.data, width, tl, dist.ctypes.data if dist is not None else None))
        return (out, dist) if want_dist else out

    @staticmethod
    def _new_records(n: int, width: int = 7) -> np.ndarray:
        """The (N,7) float64 result array.  Page-locked by default (torch's caching host allocator, so repeated calls reuse
        the block): a fresh pageable array costs more in first-touch page faults than the labelling itself (10M points:
        27 ms pageable, 16 ms page-locked, profiles/r01i_host_pipeline.md).  Bounded: arrays above TM_PINNED_OUT_MAX_MB
        (default 4096) or beyond TM_PINNED_OUT_TOTAL_MB of live results (default 8192) are ordinary np.empty arrays, as is
        everything with TM_PINNED_OUT=0."""
        global _pinned_live
        nbytes = n * 8 * width
        if (n > 0 and os.environ.get("TM_PINNED_OUT", "1") != "0"
                and nbytes <= int(os.environ.get("TM_PINNED_OUT_MAX_MB", "4096")) << 20
                and _pinned_live + nbytes <= int(os.environ.get("TM_PINNED_OUT_TOTAL_MB", "8192")) << 20):
            try:
                t = torch.empty((n, width), dtype=torch.float64, pin_memory=True)
            except RuntimeError:
                t = None
            if t is not None:
                out = t.numpy()
                _pinned_live += nbytes
                weakref.finalize(out.base, _pinned_released, nbytes)      # the tensor object numpy keeps as the array's base
                return out
        return np.empty((n, width), dtype=np.float64)

    # -- small-table fast path (QSMFittingDepthFirst.py:1006-1094) ---------------------------------
    def upload_cloud(self, cloud: np.ndarray) -> None:
        """Make a host cloud resident on the device (fp32 xyz) for ``proximity_flags``."""
        cloud = np.asarray(cloud)
        if cloud.ndim != 2 or cloud.shape[1] < 3:
            raise ValueError(f"cloud must be (N, >=3), got {cloud.shape}")
        if cloud.dtype not in (np.float32, np.float64):
            cloud = cloud.astype(np.float64)
        if cloud.shape[0] > 1 and (cloud.strides[1] != cloud.itemsize or cloud.strides[0] % cloud.itemsize
                                   or cloud.strides[0] < 3 * cloud.itemsize):
            cloud = np.ascontiguousarray(cloud)
        n = cloud.shape[0]
        row_stride = cloud.strides[0] // cloud.itemsize if n > 1 else max(3, cloud.shape[1])
        self._check(self._lib.tm_cloud_upload_host(
            self._h, cloud.ctypes.data, B.TM_F32 if cloud.dtype == np.float32 else B.TM_F64, n, row_stride))
        self._resident_n = n
        self.installs_cloud += 1

    def proximity_flags(self, subset, start, end, radius, eps: float, variant: Variant = VARIANT_B,
                        axis_eps: float = 0.0, norm_fma: bool = True, want_dist: bool = False, want_index: bool = False):
        """``distance to the closest of the given RAW cylinders < eps`` for rows ``subset`` of the resident cloud
        (``None``: every row).  Returns ``flags`` (bool) and, on request, distances and winning rows."""
        start = np.ascontiguousarray(np.asarray(start, dtype=np.float32).reshape(-1, 3))
        end = np.ascontiguousarray(np.asarray(end, dtype=np.float32).reshape(-1, 3))
        radius = np.ascontiguousarray(np.asarray(radius, dtype=np.float32).reshape(-1))
        m = start.shape[0]
        if end.shape[0] != m or radius.shape[0] != m:
            raise ValueError("cylinder arrays disagree on M")
        if subset is None:
            n, sub_ptr = getattr(self, "_resident_n", 0), None
        else:
            subset = np.ascontiguousarray(np.asarray(subset, dtype=np.int64).reshape(-1))
            n, sub_ptr = subset.shape[0], subset.ctypes.data
        flags = np.empty(n, dtype=np.uint8)
        dist = np.empty(n, dtype=np.float32) if want_dist else None
        index = np.empty(n, dtype=np.int32) if want_index else None
        prm = self._params(variant, True, norm_fma, "brute", 0.0)
        self._check(self._lib.tm_proximity_flags_host(
            self._h, sub_ptr, n, start.ctypes.data, end.ctypes.data, radius.ctypes.data, m, ctypes.byref(prm),
            float(axis_eps), float(eps), flags.ctypes.data, dist.ctypes.data if want_dist else None,
            index.ctypes.data if want_index else None))
        out = [flags.view(np.bool_)]
        if want_dist:
            out.append(dist)
        if want_index:
            out.append(index)
        return out[0] if len(out) == 1 else tuple(out)

    # -- point-neighbourhood features (Modules/Features.py:111-172) ---------------------------------
    def _f64_points(self, points) -> torch.Tensor:
        pts = torch.as_tensor(np.ascontiguousarray(np.asarray(points)[:, :3], dtype=np.float64) if not isinstance(points, torch.Tensor)
                              else points[:, :3])
        return pts.to(device=self.device, dtype=torch.float64).contiguous()

    def knn_covariance(self, points, k: int, want_idx: bool = False):
        """np.cov of (k nearest neighbours - point) for every point → (N,3,3) float64 device tensor (+ (N,k) int32 rows)."""
        pts = self._f64_points(points)
        n = pts.shape[0]
        cov = torch.empty((n, 3, 3), dtype=torch.float64, device=self.device)
        idx = torch.empty((n, k), dtype=torch.int32, device=self.device) if want_idx else None
        self._check(self._lib.tm_knn_covariance(self._h, _ptr(pts), n, 3, int(k), _ptr(cov), _ptr(idx), _stream_ptr(self.device)))
        return (cov, idx) if want_idx else cov

    def radius_count(self, points, radius: float) -> torch.Tensor:
        """Number of points within ``radius`` of every point (itself included) → (N,) int32 device tensor."""
        pts = self._f64_points(points)
        n = pts.shape[0]
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        self._check(self._lib.tm_radius_count(self._h, _ptr(pts), n, 3, float(radius), _ptr(out), _stream_ptr(self.device)))
        return out

    def noise_cloud(self, cyl_rec: torch.Tensor, first_point: torch.Tensor, n: int | None = None, point0: int = 0, seed: int = 0,
                    variates=None, want_f32: bool = False):
        """Rows ``point0 .. point0+n`` of the noisy surface cloud of a QSM (NoiseDataGeneration.py:60-102) → (n,3) float64
        device tensor (and its float32 rounding with ``want_f32``).  ``cyl_rec`` (M,14) float64 and ``first_point`` (M+1,) int64
        device tensors as described in ``include/treemorph_nn.h``; ``variates`` = (theta, z, noise) device tensors to replay
        given draws instead of the Philox stream of ``seed``."""
        if cyl_rec.dtype != torch.float64 or cyl_rec.dim() != 2 or cyl_rec.shape[1] != 14 or not cyl_rec.is_contiguous():
            raise ValueError("cyl_rec must be a contiguous (M,14) float64 tensor")
        m = cyl_rec.shape[0]
        if first_point.dtype != torch.int64 or first_point.shape != (m + 1,) or not first_point.is_contiguous():
            raise ValueError("first_point must be a contiguous (M+1,) int64 tensor")
        if cyl_rec.device != self.device or first_point.device != self.device:
            raise ValueError(f"cyl_rec / first_point must live on {self.device}")
        if n is None:
            n = (int(first_point[-1].item()) if m else 0) - point0
        out = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        out32 = torch.empty((n, 3), dtype=torch.float32, device=self.device) if want_f32 else None
        th = z = ns = None
        if variates is not None:
            th, z, ns = (torch.as_tensor(v, dtype=torch.float64, device=self.device).contiguous() for v in variates)
            if not (th.shape == z.shape == ns.shape == (n,)):
                raise ValueError("variates must be three (n,) arrays")
        self._check(self._lib.tm_noise_cloud(self._h, _ptr(cyl_rec), _ptr(first_point), m, n, int(point0), int(seed) & (2**64 - 1),
                                             _ptr(th) if th is not None else None, _ptr(z) if z is not None else None,
                                             _ptr(ns) if ns is not None else None, _ptr(out),
                                             _ptr(out32) if out32 is not None else None, _stream_ptr(self.device)))
        return (out, out32) if want_f32 else out

    def host_pipeline_info(self) -> dict:
        """D2H bytes per point and host worker threads of the last ``label_cloud_host`` call."""
        b, t = ctypes.c_int32(), ctypes.c_int32()
        self._check(self._lib.tm_host_pipeline_info(self._h, ctypes.byref(b), ctypes.byref(t)))
        return {"d2h_bytes_per_point": int(b.value), "host_threads": int(t.value)}

    def host_bandwidth(self) -> dict:
        """Copy bandwidth (read + written bytes per second) of the host memory system with the library's worker threads."""
        v, t = ctypes.c_double(), ctypes.c_int32()
        self._check(self._lib.tm_measure_host_bandwidth(self._h, ctypes.byref(v), ctypes.byref(t)))
        return {"bytes_per_s": float(v.value), "threads": int(t.value)}

    # -- introspection -----------------------------------------------------------------------
    def stats(self) -> dict:
        s = B.TmStats()
        self._check(self._lib.tm_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def set_profiling(self, enabled: bool) -> None:
        self._check(self._lib.tm_set_profiling(self._h, int(bool(enabled))))

    def phase_ms(self) -> dict:
        """Device time of each phase of the last ``label`` call (needs ``set_profiling(True)``)."""
        buf = (ctypes.c_float * B.TM_PHASES)()
        self._check(self._lib.tm_get_phase_ms(self._h, buf))
        return dict(zip(B.PHASE_NAMES, (float(v) for v in buf)))

    def selftest_arithmetic(self, n: int = 1 << 24, seed: int = 1) -> int:
        """Mismatches of the kernels' division / square-root sequences against IEEE __fdiv_rn / __fsqrt_rn."""
        bad = ctypes.c_uint64()
        self._check(self._lib.tm_selftest_arithmetic(self._h, int(n), int(seed), ctypes.byref(bad)))
        return int(bad.value)

    def fp32_peak(self) -> float:
        """Measured FP32 lane-operations per second of this GPU (FFMA / FADD+FMUL chains)."""
        v = ctypes.c_double()
        self._check(self._lib.tm_measure_fp32_peak(self._h, ctypes.byref(v)))
        return float(v.value)


def comm_init_all(engines: list[Engine]) -> None:
    """One process driving several devices: one communicator over the engines' devices (tm_comm_init_all)."""
    lib = B.load()
    arr = (ctypes.c_void_p * len(engines))(*[e._h for e in engines])
    B.check(lib, engines[0]._h, lib.tm_comm_init_all(arr, len(engines)))


def broadcast_cylinders_all(engines: list[Engine], start, radius, axis_length, axis_unit, ids=None, root_index: int = 0) -> None:
    """Replicate the table held on engines[root_index]'s device on every engine of the communicator and install it
    (tm_broadcast_cylinders_all)."""
    lib = B.load()
    root = engines[root_index]
    start = root._f32(start, 3)
    unit = root._f32(axis_unit, 3)
    m = start.shape[0]
    length = root._f32(axis_length).reshape(-1)
    radius = root._f32(radius).reshape(-1)
    if ids is not None:
        ids = torch.as_tensor(ids).to(device=root.device, dtype=torch.int32).reshape(-1)
    torch.cuda.synchronize(root.device)
    arr = (ctypes.c_void_p * len(engines))(*[e._h for e in engines])
    B.check(lib, root._h, lib.tm_broadcast_cylinders_all(
        arr, len(engines), _ptr(start), start.stride(0), start.stride(1), _ptr(unit), unit.stride(0), unit.stride(1),
        _ptr(length), length.stride(0) if m else 1, _ptr(radius), radius.stride(0) if m else 1,
        _ptr(ids), (ids.stride(0) if m else 1) if ids is not None else 1, m, int(root_index)))
    for e in engines:
        e.m = m
        e.installs += 1


_engines: dict[int, Engine] = {}


def get_engine(device: torch.device | int | str | None = None) -> Engine:
    """Process-wide engine per device (the drop-in modules share it)."""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    eng = _engines.get(idx)
    if eng is None or eng._h is None:
        eng = Engine(torch.device("cuda", idx))
        _engines[idx] = eng
    return eng
