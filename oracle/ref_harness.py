"""TEST INFRASTRUCTURE — loader for the *unmodified* reference implementation.

Only usable in the build container, where the reference checkout is mounted read-only at
``/root/reference``; the GPU box has no such directory, so nothing in ``-m gpu`` tests,
``smoke()`` or ``bench.py`` imports this module.  It is used by

* ``tests/golden/make_golden.py``  – generates the committed golden vectors, and
* ``tests/test_oracle_vs_reference.py`` – pins the C/numpy oracle against the reference
  (skipped when ``/root/reference`` is absent).

The reference needs ``fastprogress`` (absent from this image); a four-line stub of the two
callables it uses is injected into ``sys.modules``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TREEMORPH_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "PreProcessing", "LabelGenerationCuda.py"))


def _stub_fastprogress() -> None:
    if "fastprogress" in sys.modules:
        return
    fp = types.ModuleType("fastprogress")
    fp.progress_bar = lambda it, parent=None, **kw: it
    fp.master_bar = lambda it, **kw: it
    sub = types.ModuleType("fastprogress.fastprogress")
    sub.progress_bar, sub.master_bar = fp.progress_bar, fp.master_bar
    fp.fastprogress = sub
    sys.modules["fastprogress"] = fp
    sys.modules["fastprogress.fastprogress"] = sub


_cache: dict[str, types.ModuleType] = {}


def load(variant: str) -> types.ModuleType:
    """``variant`` 'A' → PreProcessing/LabelGenerationCuda.py, 'B' → Modules/Projection.py."""
    if variant in _cache:
        return _cache[variant]
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _stub_fastprogress()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if variant == "A":
        path = os.path.join(REFERENCE_ROOT, "PreProcessing", "LabelGenerationCuda.py")
        spec = importlib.util.spec_from_file_location("reference_labelgen", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    elif variant == "B":
        # the reference's package is also called ``Modules``; make sure ours is not shadowing it
        for name in [k for k in sys.modules if k == "Modules" or k.startswith("Modules.")]:
            if REFERENCE_ROOT not in (getattr(sys.modules[name], "__file__", "") or ""):
                del sys.modules[name]
        mod = importlib.import_module("Modules.Projection")
    else:
        raise ValueError(variant)
    _cache[variant] = mod
    return mod


def run_kernel(variant: str, points, start, radius, axis_length, axis_unit, ids):
    """One call of the reference's ``closest_cylinder_cuda_batch`` on CPU tensors."""
    import torch
    mod = load(variant)
    dev = torch.device("cpu")
    return mod.closest_cylinder_cuda_batch(points, start, radius, axis_length, axis_unit, ids, dev)


def run_cloud(variant: str, cloud, cylinders_df, batch_size: int = 1024):
    """The reference's ``generate_offset_cloud_cuda_batched`` on CPU → (N,7) float64."""
    import torch
    mod = load(variant)
    return mod.generate_offset_cloud_cuda_batched(cloud, cylinders_df, torch.device("cpu"),
                                                  batch_size=batch_size)
