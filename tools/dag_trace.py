"""Persistent tile-DAG Cholesky: correctness against LAPACK and a per-task trace of one factorisation.

    SPW_DAG_TRACE=1 python tools/dag_trace.py [N] [extra_rows 0|1]

Prints the wall time per factorisation (through spw_chol / a short hlm chain), the error against scipy's dpotrf,
and from the device trace: time span, chain steps (duration, gap to the previous step), per-type task durations and
the busy fraction of the sub-workers."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import spwombling_b200 as spw
from spwombling_b200._lib import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
check = os.environ.get("DAG_CHECK", "1") == "1"
lib = load()
rng = np.random.Generator(np.random.Philox(1))
coords = rng.random((n, 2))
y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)

if check:
    import oracle
    A = oracle.cov_matern2(oracle.dist_matrix(coords), 3.0, 1.5) + 0.7 * np.eye(n)
    t0 = time.perf_counter()
    U, half = spw.chol(A)
    t1 = time.perf_counter()
    import scipy.linalg as sla
    Uo = sla.cholesky(A, lower=False)
    err = np.max(np.abs(U - Uo)) / np.max(np.abs(Uo))
    print("chol n=%d: max rel err vs dpotrf %.2e, half logdet %.12g vs %.12g  (%.1f ms incl. copies)"
          % (n, err, half, np.log(np.diag(Uo)).sum(), (t1 - t0) * 1e3), flush=True)
    assert err < 1e-11, err

niter = int(os.environ.get("DAG_NITER", "6"))
r6 = np.random.Generator(np.random.Philox(6))
norm, unif = r6.standard_normal(3 * niter), r6.random(3 * niter)
kw = dict(y=y, coords=coords, nburn=0, report=max(1, niter // 2), cov_type="matern2", verbose=False)
spw.hlmBayes_sp(niter=2, norm=norm, unif=unif, **kw)
best = 1e30
for rep in range(3):
    t0 = time.perf_counter()
    spw.hlmBayes_sp(niter=niter, norm=norm, unif=unif, **kw)
    best = min(best, time.perf_counter() - t0)
print("hlm n=%d: %.2f it/s, %.3f ms per factorisation, %.2f TFLOP/s in Cholesky"
      % (n, niter / best, best / niter / 3 * 1e3, niter * n ** 3 / best / 1e12), flush=True)

cap = 4 * 120000
buf = (ctypes.c_ulonglong * cap)()
cnt = lib.spw_debug_dag_trace(n, 1, ctypes.cast(buf, ctypes.c_void_p), cap)
if cnt <= 0:
    print("no trace (set SPW_DAG_TRACE=1)")
    sys.exit(0)
tr = np.frombuffer(buf, dtype=np.uint64, count=cnt).reshape(-1, 4).astype(np.int64)
ids, t0s, t1s, who = tr[:, 0], tr[:, 1], tr[:, 2], tr[:, 3]
typ = ids >> 48
cb = (ids >> 32) & 0xffff
rb = (ids >> 16) & 0xffff
nk = ids & 0xff

start = t0s.min()
span = (t1s.max() - start) / 1e3
print("trace: %d tasks, span %.1f us" % (len(tr), span))
ch = np.where(typ == 2)[0]
ch = ch[np.argsort(t0s[ch])]
d = (t1s[ch] - t0s[ch]) / 1e3
gap = (t0s[ch][1:] - t1s[ch][:-1]) / 1e3
print("chain: %d steps, factor %.2f us mean (sum %.1f us), wait before a step %.2f us mean (sum %.1f us), max %.1f"
      % (len(ch), d.mean(), d.sum(), gap.mean(), gap.sum(), gap.max()))
print("  waits by step (us):", np.round(gap[:12], 1), "...", np.round(gap[-8:], 1))
w3 = who[ch]
sw_end, bg_seen, bg_end = (w3 & 0xfffff) * 0.016, ((w3 >> 20) & 0xfffff) * 0.016, ((w3 >> 40) & 0xfffff) * 0.016
if bg_end.any():
    print("  band mode, from the start of a step (us): sweep published %.2f mean | background sees its flags %.2f mean "
          "(p95 %.2f) | background done %.2f mean" % (sw_end.mean(), bg_seen.mean(), np.percentile(bg_seen, 95), bg_end.mean()))
    print("  background flag wait by step:", np.round(bg_seen[:10], 1), "...", np.round(bg_seen[30:40], 1), "...", np.round(bg_seen[-6:], 1))
narrow = ((ids >> 15) & 1) == 1
for name, sel in (("panel", (typ == 0) & ~narrow), ("panel 64-row", (typ == 0) & narrow),
                  ("update K=64", (typ == 1) & (nk == 1) & ~narrow), ("update K=64 64-row", (typ == 1) & (nk == 1) & narrow),
                  ("update bulk (K=64*nk)", (typ == 1) & (nk > 1))):
    if sel.any():
        dd = (t1s[sel] - t0s[sel]) / 1e3
        print("%-13s %6d tasks, mean %.2f us, p50 %.2f, p95 %.2f, total %.1f ms-SMhalf" % (name, sel.sum(), dd.mean(), np.median(dd),
              np.percentile(dd, 95), dd.sum() / 1e3))
wk = typ != 2
busy = (t1s[wk] - t0s[wk]).sum() / 1e3
nw = len(np.unique(who[wk]))
print("sub-workers: %d, busy %.1f%% of span" % (nw, 100 * busy / (nw * span)))
# utilisation over time (10 slices)
edges = np.linspace(0, span, 11)
for i in range(10):
    lo, hi = start + edges[i] * 1e3, start + edges[i + 1] * 1e3
    ov = np.clip(np.minimum(t1s[wk], hi) - np.maximum(t0s[wk], lo), 0, None).sum()
    nch = ((t1s[ch] > lo) & (t1s[ch] <= hi)).sum()
    print("  slice %d: workers busy %.0f%%, chain steps finished %d" % (i, 100 * ov / (nw * (hi - lo)), nch))
if os.environ.get("DAG_DUMP"):
    np.save(os.environ["DAG_DUMP"], tr)
