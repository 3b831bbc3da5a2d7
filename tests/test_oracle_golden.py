"""The CPU oracle against the golden vectors minted from the reference itself (bit-for-bit).

The fixtures were produced by ``tests/golden/make_golden.py`` executing
``PreProcessing/LabelGenerationCuda.py`` (variant A) and ``Modules/Projection.py`` (variant B)
on CPU; see SURVEY.md §8(c).  This is what pins the oracle.
"""
import numpy as np
import pytest

from conftest import assert_same_bits
from oracle import oracle


@pytest.mark.parametrize("vn", ["A", "B"])
def test_c_oracle_cloud_matches_reference(golden, vn):
    var = oracle.VARIANTS[vn]
    out = oracle.label_cloud(golden["cloud"], golden["qsm"], var)
    assert out.dtype == np.float64 and out.shape == (len(golden["cloud"]), 7)
    assert_same_bits(out, golden[f"ref_{vn}_cloud"], f"{golden['name']} variant {vn} (N,7)")


@pytest.mark.parametrize("vn", ["A", "B"])
def test_c_oracle_kernel_matches_reference(golden, vn):
    var = oracle.VARIANTS[vn]
    q = golden["qsm"]
    start = np.stack([q["startX"], q["startY"], q["startZ"]], 1).astype(np.float32)
    end = np.stack([q["endX"], q["endY"], q["endZ"]], 1).astype(np.float32)
    length, unit = oracle.prepare(start, end, var, norm_fma=(len(start) == 1))
    assert_same_bits(length, golden[f"ref_{vn}_length"], "axis_length")
    assert_same_bits(unit, golden[f"ref_{vn}_unit"], "axis_unit")
    res = oracle.label(golden["cloud"][:, :3], start, q["radius"], length, unit, q["ID"], var)
    assert_same_bits(res["id"], golden[f"ref_{vn}_id"], "ids")
    assert_same_bits(res["dist"], golden[f"ref_{vn}_dist"], "distances")
    assert_same_bits(res["offset"], golden[f"ref_{vn}_off"], "offsets")


@pytest.mark.parametrize("vn", ["A", "B"])
def test_numpy_oracle_matches_reference(golden, vn):
    var = oracle.VARIANTS[vn]
    res = oracle.label_numpy(golden["cloud"][:400, :3],
                             np.stack([golden["qsm"][k] for k in ("startX", "startY", "startZ")], 1),
                             golden["qsm"]["radius"], golden[f"ref_{vn}_length"], golden[f"ref_{vn}_unit"],
                             golden["qsm"]["ID"], var)
    assert_same_bits(res["id"], golden[f"ref_{vn}_id"][:400], "ids")
    assert_same_bits(res["dist"], golden[f"ref_{vn}_dist"][:400], "distances")
    assert_same_bits(res["offset"], golden[f"ref_{vn}_off"][:400], "offsets")


def test_edge_semantics_documented_in_survey_a4():
    """Duplicate cylinders → lowest row; on-axis point: A → NaN wins, B → distance 0."""
    from conftest import load_golden
    g = load_golden("adversarial")
    a_id, b_id = g["ref_A_id"], g["ref_B_id"]
    assert a_id[0] == 5 and b_id[0] == 5                     # tie between rows 0 and 1 → row 0 (ID 5)
    assert np.isnan(g["ref_A_dist"][1]) and a_id[1] == 5     # on the axis: NaN beats finite
    assert g["ref_B_dist"][1] == 0.0
    z = load_golden("zero_length")
    assert np.isnan(z["ref_A_dist"]).all()                   # zero-length cylinder poisons variant A
    assert np.isfinite(z["ref_B_dist"]).all()


def test_closed_form_agrees_with_oracle():
    """Independent derivation in float64 vs the mirror-order fp32 oracle."""
    from conftest import load_golden
    g = load_golden("tree300")
    for vn in "AB":
        var = oracle.VARIANTS[vn]
        q = g["qsm"]
        start = np.stack([q["startX"], q["startY"], q["startZ"]], 1).astype(np.float32)
        d32 = oracle.distance_matrix(g["cloud"][:200, :3], start, q["radius"], g[f"ref_{vn}_length"],
                                     g[f"ref_{vn}_unit"], var)
        d64 = oracle.closed_form_f64(g["cloud"][:200, :3], start, q["radius"], g[f"ref_{vn}_length"],
                                     g[f"ref_{vn}_unit"], var)
        # the perp switch is decided on a rounded d, so exclude the 1e-6/1e-3 boundary band from the comparison
        assert np.nanmax(np.abs(d32 - d64)) < 2e-3
        assert np.nanmedian(np.abs(d32 - d64)) < 1e-6


def test_m_zero_raises_like_the_reference():
    with pytest.raises(IndexError):
        oracle.label(np.zeros((3, 3), np.float32), np.zeros((0, 3)), np.zeros(0), np.zeros((0, 1)), np.zeros((0, 3)))
    out = oracle.label(np.zeros((0, 3), np.float32), np.zeros((1, 3)), np.ones(1), np.ones((1, 1)), np.ones((1, 3)))
    assert out["id"].shape == (0,) and out["offset"].shape == (0, 3)
