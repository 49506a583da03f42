"""Per-step trace of the look-ahead Cholesky: SPW_LA_TRACE=1 python tools/la_trace.py N [reps]
For every step j (times in us relative to the moment the chain published step j-2, i.e. when LLS(j)'s inputs exist):
  D wait / D step, PANEL2(j-2) start..end, LLS(j) start..end."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spwombling_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ms = ctypes.c_double(0); info = ctypes.c_int(0)
rc = lib.spw_debug_potrf_time(n, 1, reps, ctypes.byref(ms), ctypes.byref(info))
print("rc=%d status=%d %.3f ms" % (rc, info.value, ms.value))
nblk = (n + 63) // 64
buf = np.zeros(8 * nblk, dtype=np.uint64)
cnt = lib.spw_debug_la_trace(n, 1, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
if cnt <= 0:
    sys.exit("no trace (SPW_LA_TRACE=1?)")
t = buf.reshape(nblk, 8).astype(np.float64) / 1e3
t0 = t[0, 0]
wait = t[:, 1] - t[:, 0]
step = t[:, 2] - t[:, 1]
print("span %.1f us; steps: sum %.1f us (mean %.2f); waits: sum %.1f us" % (t[-1, 2] - t0, step.sum(), step.mean(), wait.sum()))
lo, hi = int(os.environ.get("LA_TRACE_FROM", "0")), int(os.environ.get("LA_TRACE_TO", str(nblk)))
print("  j   t(us)  Dwait Dstep (look-ahead + sweep) | since D(j-2) published: P2(j-2) start  end | LLS(j) start   end | D(j) starts")
for j in range(max(2, lo), min(hi, nblk)):
    ref = t[j - 2, 2]
    print("%3d %7.1f %6.1f %5.1f (%4.1f + %4.1f) | %28.1f %5.1f | %12.1f %5.1f | %6.1f" % (
        j, t[j, 0] - t0, wait[j], step[j], t[j, 7] - t[j, 1], t[j, 2] - t[j, 7], t[j - 2, 3] - ref, t[j - 2, 4] - ref, t[j, 5] - ref, t[j, 6] - ref, t[j, 1] - ref))
