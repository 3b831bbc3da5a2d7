"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the binding covers the same set, and the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

from treemorph_b200 import binding, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "treemorph_nn.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    path = build.build()
    assert os.path.dirname(path) == build.PKG_DIR and os.path.exists(path)


def test_every_header_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 10
    lib = ctypes.CDLL(build.build())
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in binding.SIGNATURES, f"{name} declared in the header but not bound"
    assert sorted(binding.SIGNATURES) == names


def test_abi_version_and_struct_layout():
    lib = binding.load()
    assert lib.tm_version() == binding.ABI_VERSION
    assert ctypes.sizeof(binding.TmParams) == 32
    assert ctypes.sizeof(binding.TmStats) == 128
    assert lib.tm_status_string(binding.TM_ERR_NO_CYLINDERS).decode().startswith("argmin()")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = binding.load()
    h = ctypes.c_void_p()
    assert lib.tm_create(0, ctypes.byref(h)) == binding.TM_ERR_CUDA and not h.value
    from treemorph_b200 import api
    with pytest.raises(RuntimeError):
        api.Engine()
    from treemorph_b200.PreProcessing import LabelGenerationCuda as L
    import numpy as np
    with pytest.raises(RuntimeError):
        L.closest_cylinder_cuda_batch(np.zeros((2, 3), np.float32), torch.zeros(1, 3), torch.ones(1), torch.ones(1, 1),
                                      torch.ones(1, 3), torch.zeros(1, dtype=torch.int32), torch.device("cpu"))


def test_product_path_does_not_import_the_oracle():
    pkg = build.PKG_DIR
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/nearest_cylinder.c", ""), f"{f} mentions the oracle"
