"""Boundary helpers of the hot path with the reference's names (``Modules/Utils.py:146-158, 190-250``).

``get_device`` differs from the reference in one deliberate way: the reference silently falls back to
``torch.device('cpu')``; this build has no CPU path for the nearest-cylinder search, so the absence of a
CUDA device is an error (asked for explicitly with ``GPU=False`` it still returns the cpu device, which
the labelling functions then reject).
"""
from __future__ import annotations

import os

import numpy as np
import torch

try:                                    # optional, exactly as in the reference (Utils.py:182-187)
    import laspy
    HAS_LASPY = True
except ImportError:                     # pragma: no cover - laspy is absent from this image
    laspy = None
    HAS_LASPY = False


def get_device(GPU=True):
    if GPU:
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the B200 nearest-cylinder path has no CPU fallback")
        idx = torch.cuda.current_device()
        print("Using cuda device")
        print(f"Using CUDA Device: {torch.cuda.get_device_name(idx)}")
        return torch.device("cuda")
    print("Using cpu")
    return torch.device("cpu")


def _read_table(path):
    last = None
    for delim in (" ", ","):
        try:
            return np.loadtxt(path, delimiter=delim)
        except ValueError as exc:
            last = exc
    print(f"ERROR: Could not parse TXT {path} with space or comma delimiter. ({last})")
    return None


def load_cloud(path):
    """XYZ of a ``.npy`` / ``.txt`` / ``.las`` / ``.laz`` cloud as float32 ``(N,3)``; ``None`` on any failure."""
    ext = os.path.splitext(path)[1].lower()
    try:
        if ext == ".npy":
            pts = np.load(path)
            if pts.ndim == 1:
                if pts.size % 3:
                    print(f"ERROR: .npy file {path} is 1D and not reshapeable to (N,3). Shape: {pts.shape}")
                    return None
                pts = pts.reshape(-1, 3)
        elif ext == ".txt":
            pts = _read_table(path)
        elif ext in (".las", ".laz"):
            if not HAS_LASPY:
                print(f"ERROR: Cannot load {path}. laspy is not installed or import failed.")
                return None
            with laspy.open(path) as fh:
                las = fh.read()
            pts = np.vstack((las.x, las.y, las.z)).T
        else:
            print(f"ERROR: Unsupported file format: {ext} for {path}")
            return None
    except FileNotFoundError:
        print(f"ERROR: File not found: {path}")
        return None
    except Exception as exc:            # same contract as the reference: report and return None
        print(f"ERROR: Failed to load point cloud from {path}: {exc}")
        return None
    if pts is None:
        return None
    if pts.ndim != 2 or pts.shape[1] < 3:
        print(f"ERROR: Loaded data from {path} has unexpected shape after processing: {pts.shape}. Expected (N, >=3).")
        return None
    return pts[:, :3].astype(np.float32)
