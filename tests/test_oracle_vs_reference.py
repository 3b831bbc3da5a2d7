"""Pins the oracle against the live reference (build container only; skipped on the GPU box,
where /root/reference does not exist).  Larger and more random than the committed fixtures."""
import numpy as np
import pytest

from conftest import assert_same_bits
from oracle import oracle, ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not mounted")


def _synth():
    import importlib
    return importlib.import_module("treemorph_b200.synth")


@pytest.mark.parametrize("vn", ["A", "B"])
def test_dataframe_path_bit_exact(vn):
    synth = _synth()
    q = synth.random_qsm(600, seed=101)
    pts = synth.sample_points(q, 4096 + 77, seed=102)          # ragged last batch
    ref = ref_harness.run_cloud(vn, pts, synth.qsm_dataframe(q), batch_size=1024)
    out = oracle.label_cloud(pts, q, oracle.VARIANTS[vn])
    assert_same_bits(out, ref, f"variant {vn}")


@pytest.mark.parametrize("vn", ["A", "B"])
def test_contiguous_tensor_path_bit_exact(vn):
    """C-ordered cylinder tensors (as QSMFittingDepthFirst.py:1039-1045 builds them) take ATen's
    contiguous norm path → the oracle's norm_fma=True rounding."""
    import torch
    synth = _synth()
    var = oracle.VARIANTS[vn]
    q = synth.random_qsm(40, seed=111)
    pts = synth.sample_points(q, 3000, seed=112)
    start, radius, length, unit, ids = synth.cylinder_arrays(q, var.axis_eps)
    rid, rdist, roff = ref_harness.run_kernel(vn, pts, torch.tensor(start), torch.tensor(radius),
                                              torch.tensor(length), torch.tensor(unit), torch.tensor(ids))
    o = oracle.label(pts, start, radius, length, unit, ids, var, norm_fma=True)
    assert_same_bits(o["id"], rid, "ids")
    assert_same_bits(o["dist"], rdist, "distances")
    assert_same_bits(o["offset"], roff, "offsets")


@pytest.mark.parametrize("vn", ["A", "B"])
def test_torch_mirror_is_the_reference_chain(vn):
    """oracle/torch_mirror.py (the comparator that runs on the GPU box, where the reference cannot travel) issues the
    reference's own ATen calls: on the CPU it must reproduce the live reference bit for bit."""
    import torch
    from oracle import torch_mirror
    synth = _synth()
    var = oracle.VARIANTS[vn]
    q = synth.random_qsm(500, seed=121)
    pts = synth.sample_points(q, 2048 + 5, seed=122)
    ref = ref_harness.run_cloud(vn, pts, synth.qsm_dataframe(q), batch_size=1024)
    got = torch_mirror.label(pts, q, torch.device("cpu"), var.perp_atol, var.norm_eps, var.axis_eps, fortran=True)
    assert_same_bits(got["offset"].astype(np.float64), ref[:, 3:6], f"variant {vn}: offsets")
    assert_same_bits(got["id"].astype(np.float64), ref[:, 6], f"variant {vn}: ids")
