"""Per-step trace of the look-ahead Cholesky's diagonal chain: SPW_LA=1 SPW_LA_TRACE=1 python tools/la_trace.py N"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spwombling_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
ms = ctypes.c_double(0); info = ctypes.c_int(0)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rc = lib.spw_debug_potrf_time(n, 1, reps, ctypes.byref(ms), ctypes.byref(info))
print("rc=%d status=%d %.3f ms" % (rc, info.value, ms.value))
nblk = (n + 63) // 64
buf = np.zeros(3 * nblk, dtype=np.uint64)
cnt = lib.spw_debug_la_trace(n, 1, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
if cnt <= 0:
    sys.exit("no trace (SPW_LA=1 SPW_LA_TRACE=1?)")
t = buf.reshape(nblk, 3).astype(np.int64)
t0 = t[0, 0]
wait = (t[:, 1] - t[:, 0]) / 1e3
step = (t[:, 2] - t[:, 1]) / 1e3
print("span %.1f us; steps: sum %.1f us (mean %.2f); waits: sum %.1f us" % ((t[-1, 2] - t0) / 1e3, step.sum(), step.mean(), wait.sum()))
nshow = int(os.environ.get("LA_TRACE_ROWS", "100"))
for m in range(0, min(nblk, 4 * nshow), 4):
    sl = slice(m, min(m + 4, nblk))
    print("block %2d  t=%7.1f us  step %s  wait %s" % (m // 4, (t[m, 0] - t0) / 1e3,
          " ".join("%5.1f" % v for v in step[sl]), " ".join("%6.1f" % v for v in wait[sl])))
