"""Debug helper: one grid-mode labelling of a seeded case, compared with brute mode (run under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from treemorph_b200 import api, synth

m = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
vn = sys.argv[3] if len(sys.argv) > 3 else "A"
q = synth.random_qsm(m, seed=1)
pts = synth.sample_points(q, n, seed=2)
var = api.VARIANTS[vn]
start, radius, length, unit, ids = synth.cylinder_arrays(q, var.axis_eps)
eng = api.Engine()
dev = eng.device
eng.set_cylinders(*(torch.tensor(x, device=dev) for x in (start, radius, length, unit)), torch.tensor(ids, device=dev))
d = torch.tensor(pts, device=dev)
g = eng.label(d, var, mode="grid", want=("index", "id", "dist", "offset"))
torch.cuda.synchronize()
print("grid ok", eng.stats())
b = eng.label(d, var, mode="brute", want=("index", "id", "dist", "offset"))
torch.cuda.synchronize()
for k in g:
    same = (g[k] == b[k]) | ((g[k] != g[k]) & (b[k] != b[k])) if g[k].dtype.is_floating_point else g[k] == b[k]
    print(k, "mismatches:", int((~same).sum()))
