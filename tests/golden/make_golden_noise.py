"""Golden vectors for the noisy-cloud generator, produced by running the UNMODIFIED reference
(``PreProcessing/NoiseDataGeneration.py:14-106``) on QSM files written to a temporary directory.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_noise.py

The reference draws from numpy's global legacy generator; each case seeds it (``np.random.seed``) right before the call,
so the variates can be regenerated in the tests from the seed alone.  Stored in noise.npz per case: the CSV text, the file
name, the seed and the (N,3) float64 cloud the reference saved.
"""
from __future__ import annotations

import importlib.util
import io
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "extracting-tree-morphology-from-point-clouds_b200"))

import ref_harness  # noqa: E402
import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    ref_harness._stub_fastprogress()
    if "matplotlib" not in sys.modules:                     # imported at module level, never used by the generator
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    path = os.path.join(ref_harness.REFERENCE_ROOT, "PreProcessing", "NoiseDataGeneration.py")
    spec = importlib.util.spec_from_file_location("reference_noisegen", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def csv_text(qsm: dict, padded: bool = False) -> str:
    import pandas as pd
    df = pd.DataFrame(qsm)
    if padded:
        df.columns = [" " + c for c in df.columns]          # the generator strips header whitespace (:31)
    buf = io.StringIO()
    df.to_csv(buf, index=False)
    return buf.getvalue()


def cases():
    tree = synth.random_qsm(60, seed=11)
    yield "tree60", "33_22_000000.csv", csv_text(tree), 1234
    # axis-aligned cylinders (+z: the v = (1,0,0) substitution of :83), a tilted one, a thin twig that gets no points
    qsm = {
        "startX": [0.0, 1.0, 0.5, 2.0], "startY": [0.0, 0.0, 0.5, 2.0], "startZ": [0.0, 0.0, 1.0, 1.5],
        "endX": [0.0, 1.0, 0.9, 2.05], "endY": [0.0, 0.0, 0.2, 2.0], "endZ": [1.0, 0.6, 1.7, 1.6],
        "radius": [0.20, 0.12, 0.08, 0.001], "ID": [7, 3, 9, 4],
    }
    yield "aligned", "plot_7.csv", csv_text({k: np.array(v) for k, v in qsm.items()}, padded=True), 99
    big = synth.random_qsm(400, seed=12)
    yield "tree400", "1_2_x_y.csv", csv_text(big), 2026


def main():
    ref = load_reference()
    rec = {}
    names = []
    for name, fname, text, seed in cases():
        with tempfile.TemporaryDirectory() as tmp:
            src, dst = os.path.join(tmp, "qsm"), os.path.join(tmp, "cloud")
            os.makedirs(src)
            os.makedirs(dst)
            with open(os.path.join(src, fname), "w") as f:
                f.write(text)
            np.random.seed(seed)
            ref.noiseGeneration(src, dst)
            written = sorted(os.listdir(dst))
            assert len(written) == 1, written
            cloud = np.load(os.path.join(dst, written[0]))
        names.append(name)
        rec[f"{name}__csv"] = np.array(text)
        rec[f"{name}__file"] = np.array(fname)
        rec[f"{name}__written"] = np.array(written[0])
        rec[f"{name}__seed"] = np.array(seed)
        rec[f"{name}__cloud"] = cloud
        print(name, fname, "->", written[0], cloud.shape, cloud.dtype)
    rec["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "noise.npz"), **rec)
    print("wrote", os.path.join(OUT, "noise.npz"), os.path.getsize(os.path.join(OUT, "noise.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
