"""pytest configuration: registers the ``gpu`` marker and puts the repo on sys.path.

``-m "not gpu"`` covers the oracle against the golden vectors, the host logic and the C-ABI
symbol table; ``-m gpu`` tests are the parity tests proper and need a B200.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


OTHER_FIXTURES = {"proximity", "features", "drivers", "noise"}      # fixtures of other entry points, with their own tests


def golden_names():
    """Golden vectors of the labelling path (tests/golden/make_golden.py)."""
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if n not in OTHER_FIXTURES)


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        rec = {k: z[k] for k in z.files}
    rec["qsm"] = {k[4:]: rec[k] for k in list(rec) if k.startswith("qsm_")}
    return rec


@pytest.fixture(params=golden_names())
def golden(request):
    rec = load_golden(request.param)
    rec["name"] = request.param
    return rec


def assert_same_bits(a, b, what=""):
    """Bit-for-bit equality that treats NaN == NaN."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = (a == b) | (np.isnan(a) & np.isnan(b)) if a.dtype.kind == "f" else (a == b)
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {a.size} elements differ"
