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


def test_implicit_tree_numbering_used_by_the_device_bvh_build():
    """csrc/tm_bvh.cu builds the hierarchy without recursion: node = binary-heap slot, range of a slot = follow the bits of
    the slot number from the root splitting count -> (count // 2, count - count // 2), leaf when count <= 4, and the slot
    array is sized 2 << depth with depth the first level where ceil(n / 2^depth) <= 4.  This mirrors those formulas in
    Python and checks them against the plain recursive split for every table size up to 3000 and a few large ones."""
    LEAF = 4

    def slot_range(slot, n):
        first, count = 0, n
        depth = slot.bit_length() - 1
        for b in range(depth - 1, -1, -1):
            if count <= LEAF:
                return None                                  # an ancestor is a leaf
            half = count // 2
            if (slot >> b) & 1:
                first, count = first + half, count - half
            else:
                count = half
        return first, count

    def recurse(slot, first, count, out):
        out[slot] = (first, count)
        if count > LEAF:
            half = count // 2
            recurse(2 * slot, first, half, out)
            recurse(2 * slot + 1, first + half, count - half, out)

    for n in list(range(1, 3001)) + [50_000, 199_999, 200_000, 1_000_003]:
        depth = 0
        while ((n + (1 << depth) - 1) >> depth) > LEAF:
            depth += 1
        slots = 2 << depth
        want = {}
        recurse(1, 0, n, want)
        assert max(want) < slots, n
        leaves = sorted((f, c) for f, c in want.values() if c <= LEAF)
        assert leaves[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(leaves, leaves[1:])) and sum(c for _, c in leaves) == n
        check = want if n <= 3000 else {s: want[s] for s in list(want)[:: max(1, len(want) // 500)]}
        for slot, rng in check.items():
            assert slot_range(slot, n) == rng, (n, slot)
        if n <= 300:                                          # slots that do not exist are recognised as such
            for slot in range(1, slots):
                got = slot_range(slot, n)
                assert (got is None) == (slot not in want), (n, slot)


def test_load_cloud_las_branch_with_a_stand_in_for_laspy(tmp_path, monkeypatch, capsys):
    """The ``.las`` / ``.laz`` branch of load_cloud (reference Modules/Utils.py:237-245): laspy is not in this image, so a
    stand-in with the two calls the branch makes (``laspy.open(path)`` as a context manager, ``.read()`` with ``x/y/z``) is
    injected; without it the branch reports the missing library and returns None, as the reference does."""
    from treemorph_b200.Modules import Utils
    path = tmp_path / "cloud.laz"
    path.write_bytes(b"not a real laz file")
    xyz = np.array([[1.5, 2.5, 3.5], [4.0, 5.0, 6.0], [7.25, 8.25, 9.25]])

    monkeypatch.setattr(Utils, "HAS_LASPY", False)
    assert Utils.load_cloud(str(path)) is None
    assert "laspy is not installed" in capsys.readouterr().out

    class _Reader:
        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def read(self):
            class _Las:
                x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
            return _Las()

    class _FakeLaspy:
        opened = []

        @staticmethod
        def open(p):
            _FakeLaspy.opened.append(p)
            return _Reader()

    monkeypatch.setattr(Utils, "laspy", _FakeLaspy)
    monkeypatch.setattr(Utils, "HAS_LASPY", True)
    for ext in (".laz", ".LAS"):
        p = tmp_path / ("cloud" + ext)
        p.write_bytes(b"x")
        got = Utils.load_cloud(str(p))
        assert got.dtype == np.float32 and got.shape == (3, 3) and np.array_equal(got, xyz.astype(np.float32))
    assert len(_FakeLaspy.opened) == 2

    class _Broken(_FakeLaspy):
        @staticmethod
        def open(p):
            raise OSError("truncated file")
    monkeypatch.setattr(Utils, "laspy", _Broken)
    assert Utils.load_cloud(str(path)) is None                      # any failure: message + None, like the reference
    assert "Failed to load point cloud" in capsys.readouterr().out


def test_threaded_npy_writer_produces_np_save_bytes(tmp_path):
    """dropin._write_npy (the drivers' file writer: header + payload copied into the page cache by several threads) writes
    exactly what np.save writes, for the (N,11) record, small arrays, and sizes that do not divide by the thread count."""
    from treemorph_b200 import dropin
    rng = np.random.default_rng(3)
    for shape in ((0, 11), (1, 11), (1000, 11), (300_001, 11), (1_100_000, 7)):
        a = rng.random(shape)
        dropin._write_npy(str(tmp_path / "mine.npy"), a)
        np.save(tmp_path / "ref.npy", a)
        assert (tmp_path / "mine.npy").read_bytes() == (tmp_path / "ref.npy").read_bytes(), shape
        assert np.array_equal(np.load(tmp_path / "mine.npy"), a)
