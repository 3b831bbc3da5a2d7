"""The reference's data-preparation chain through the drop-in modules, files included:

    noiseGeneration (QSM csv -> noisy cloud .npy)  ->  label_clouds (cloud + csv -> <stem>_labeled.npy, (N,11) float64)

on a handful of synthetic trees (5 000 cylinders each; the reference's own density gives ~0.25M points per tree, so the counts
are scaled to ~1M).  Wall-clock per stage; everything (pandas, np.load / np.save on the box's disk, pinned allocations after
the first tree) is inside."""
import contextlib, io, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from treemorph_b200 import synth
from treemorph_b200.PreProcessing import LabelGenerationCuda as L, NoiseDataGeneration as N

trees, target = 4, 1_000_000
with tempfile.TemporaryDirectory() as tmp:
    qdir, cdir, ldir = (os.path.join(tmp, d) for d in ("qsm", "cloud", "label"))
    for d in (qdir, cdir, ldir):
        os.makedirs(d)
    for t in range(trees):
        synth.qsm_dataframe(synth.random_qsm(5000, seed=40 + t)).to_csv(os.path.join(qdir, f"{t + 1}_{7}_000000.csv"), index=False)
    quiet = io.StringIO()
    # the reference's generator at its own density ...
    with contextlib.redirect_stdout(quiet):
        np.random.seed(3)
        t0 = time.perf_counter(); N.noiseGeneration(qdir, cdir); torch.cuda.synchronize(); first_s = time.perf_counter() - t0
        t0 = time.perf_counter(); N.noiseGeneration(qdir, cdir); torch.cuda.synchronize(); gen_s = time.perf_counter() - t0
    sizes = [len(np.load(os.path.join(cdir, f))) for f in sorted(os.listdir(cdir))]
    # ... and ~1M points per tree for the labelling stage (same generator, scaled counts)
    import pandas as pd
    for f in sorted(os.listdir(qdir)):
        df = pd.read_csv(os.path.join(qdir, f))
        plan = N.cylinder_plan(df)
        plan.counts[:] = np.maximum(1, (plan.counts * (target / plan.n_points)).astype(np.int64))
        plan.first_point[1:] = np.cumsum(plan.counts)
        np.save(os.path.join(cdir, "_".join(f.split("_")[:2]) + ".npy"), N.noise_cloud(plan, "cuda:0", seed=5))
    n_total = sum(len(np.load(os.path.join(cdir, f))) for f in os.listdir(cdir))
    out = {"trees": trees, "cylinders_per_tree": 5000, "noiseGeneration_first_call_s": first_s, "noiseGeneration_s": gen_s, "points_at_reference_density": sizes,
           "points_labelled": n_total}
    for feats in (False, True):
        for rep in range(2):
            with contextlib.redirect_stdout(quiet):
                t0 = time.perf_counter(); L.label_clouds(cdir, qdir, ldir, use_features=feats); dt = time.perf_counter() - t0
        shape = np.load(os.path.join(ldir, sorted(os.listdir(ldir))[0])).shape
        out[f"label_clouds_features_{feats}"] = {"s": dt, "s_per_tree": dt / trees, "points_per_s": n_total / dt, "record_shape": list(shape)}
    print(json.dumps(out))
