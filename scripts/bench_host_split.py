"""tm_label_cloud_host with a page-locked record array: share of the chunks assembled on the device (TM_HOST_SPLIT, percent)
versus by the host workers, and the chunk size.  10M x 50k, pinned fp32 cloud in, pinned (N,7) float64 out."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from treemorph_b200 import api, synth

dev = torch.device("cuda", 0)
eng = api.get_engine(dev)
n, m = 10_000_000, 50_000
qsm = synth.random_qsm(m, seed=3)
pts = synth.sample_points(qsm, n, seed=4, noise="model")
s, r, l, u, i = synth.cylinder_arrays(qsm)
eng.set_cylinders(*[torch.tensor(x, device=dev) for x in (s, r, l, u)], torch.tensor(i, device=dev))
cin = torch.empty((n, 3), dtype=torch.float32, pin_memory=True); cin.numpy()[:] = pts
out = torch.empty((n, 7), dtype=torch.float64, pin_memory=True)
ref = None
for depth in (2, 3, 4):
    os.environ["TM_HOST_DEPTH"] = str(depth)
    for chunk in (1 << 20, 1 << 19):
        os.environ["TM_HOST_CHUNK"] = str(chunk)
        for split in (0, 15, 20, 25, 30, 40, 100):
            os.environ["TM_HOST_SPLIT"] = str(split)
            ts = []
            for rep in range(7):
                t0 = time.perf_counter()
                eng.label_cloud_host(cin.numpy(), api.VARIANT_A, out=out.numpy())
                ts.append(time.perf_counter() - t0)
            if ref is None:
                ref = out.numpy().copy()
            same = bool(np.array_equal(ref, out.numpy(), equal_nan=True))
            t = float(np.median(ts[2:]))
            print(json.dumps({"in_flight": depth, "chunk": chunk, "split_pct": split, "ms": round(1e3 * t, 3), "gpts_per_s": round(n / t / 1e9, 3),
                              "d2h_bytes_per_point": eng.host_pipeline_info()["d2h_bytes_per_point"], "same_rows": same}), flush=True)
# the drop-in call as a user makes it: pageable float64 cloud in, the engine allocates the records
cloud64 = pts.astype(np.float64)
for k in ("TM_HOST_CHUNK", "TM_HOST_SPLIT", "TM_HOST_DEPTH"):
    os.environ.pop(k, None)
for pinned in ("0", "1"):
    os.environ["TM_PINNED_OUT"] = pinned
    ts = []
    for rep in range(5):
        t0 = time.perf_counter()
        rec = eng.label_cloud_host(cloud64, api.VARIANT_A)
        ts.append(time.perf_counter() - t0)
        del rec
    print(json.dumps({"api": "label_cloud_host(pageable f64 cloud), records allocated by the engine", "pinned_out": pinned,
                      "ms": [round(1e3 * t, 1) for t in ts]}), flush=True)
