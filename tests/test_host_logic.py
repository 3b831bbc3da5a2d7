"""Host-side logic that needs no GPU: result-array allocation policy of the Python layer."""
import numpy as np


def test_result_array_falls_back_to_pageable(monkeypatch):
    """Engine._new_records: page-locked when possible and within budget, a plain np.empty otherwise (no CUDA here, budget 0,
    switched off, or above the per-array cap) — never an error."""
    from treemorph_b200 import api
    for env in ({}, {"TM_PINNED_OUT": "0"}, {"TM_PINNED_OUT_TOTAL_MB": "0"}, {"TM_PINNED_OUT_MAX_MB": "0"}):
        for k in ("TM_PINNED_OUT", "TM_PINNED_OUT_TOTAL_MB", "TM_PINNED_OUT_MAX_MB"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = api.Engine._new_records(1000)
        assert out.shape == (1000, 7) and out.dtype == np.float64 and out.flags.c_contiguous and out.flags.writeable
    assert api.Engine._new_records(0).shape == (0, 7)
    assert api._pinned_live >= 0
