"""Scratch timing of the pieces of bench.py's e2e step (not part of the product)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import spwombling_b200 as spw
from spwombling_b200 import device as dev

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = bench.WORKLOADS[wl]
B = int(sys.argv[2]) if len(sys.argv) > 2 else w["batch"]
device = torch.device("cuda", 0)
coords, grid, phi, sig2, z, eps = bench.make_inputs(w, B)
coords_d = dev.coords_to_device(coords, device); grid_d = dev.coords_to_device(grid, device)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = dict(phi=pin(phi), sig2=pin(sig2), z=pin(z), eps=pin(eps))
c = bench.ncomp(w["cov"]); G = grid.shape[0]
buf = torch.empty((B, c, G), dtype=torch.float64, device=device)
def T(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("%-28s %8.2f ms" % (label, (time.perf_counter() - t) * 1e3), flush=True); return r
for rep in range(3):
    print("--- rep", rep)
    d = T("h2d", lambda: {k: v.to(device, non_blocking=True) for k, v in h.items()})
    mine = T("gradient_shard", lambda: dev.gradient_shard(coords_d, grid_d, w["cov"], d["phi"], d["sig2"], d["z"], d["eps"], out=buf))
    summ = T("hpd_summary_dev", lambda: dev.hpd_summary_dev(mine, 2))
    s2 = T("summary.cpu", lambda: summ.cpu())
    r = T("draws_to_r_layout", lambda: dev.draws_to_r_layout(mine))
    r2 = T("draws.cpu", lambda: r.cpu())
