"""Where does a per-tree call spend its time?  (set_cylinders + index build, then the drop-in cloud call.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from treemorph_b200 import api, synth
from treemorph_b200.Modules import Projection
dev = torch.device("cuda", 0)
eng = api.get_engine(dev)
for m, n in ((5_000, 1_000_000), (50_000, 5_000_000)):
    qsm = synth.random_qsm(m, seed=3)
    pts = synth.sample_points(qsm, n, seed=4, noise="model")
    s, r, l, u, i = synth.cylinder_arrays(qsm)
    ts = [torch.tensor(x, device=dev) for x in (s, r, l, u)] + [torch.tensor(i, device=dev)]
    dp = torch.tensor(pts[:4096], device=dev)
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.set_cylinders(*ts)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        eng.label(dp, api.VARIANT_B, mode="grid", want=("id",))
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"M={m} rep {rep}: set_cylinders {1e3*(t1-t0):.2f} ms, index build + 4096 pts {1e3*(t2-t1):.2f} ms", flush=True)
    df = synth.qsm_dataframe(qsm)
    cloud64 = pts.astype(np.float64)
    out = None
    for rep in range(4):
        t0 = time.perf_counter()
        rec = Projection.generate_offset_cloud_cuda_batched(cloud64, df, dev)
        t1 = time.perf_counter()
        print(f"M={m} N={n} rep {rep}: generate_offset_cloud_cuda_batched {1e3*(t1-t0):.1f} ms  ({n/(t1-t0)/1e6:.1f} Mpts/s)", flush=True)
    # pieces of the drop-in call
    t0 = time.perf_counter(); a = np.empty((n, 7)); a[:] = 0; t1 = time.perf_counter()
    print(f"   np.empty((N,7)) + first touch: {1e3*(t1-t0):.1f} ms")
    t0 = time.perf_counter(); eng.label_cloud_host(cloud64, api.VARIANT_B, out=a); t1 = time.perf_counter()
    print(f"   label_cloud_host into a touched array: {1e3*(t1-t0):.1f} ms")
    c32 = pts.copy()
    t0 = time.perf_counter(); eng.label_cloud_host(c32, api.VARIANT_B, out=a); t1 = time.perf_counter()
    print(f"   label_cloud_host fp32 pageable in: {1e3*(t1-t0):.1f} ms")
