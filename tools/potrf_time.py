"""Time one single-matrix factorisation under several schedules: python tools/potrf_time.py N [reps] -- each remaining
argument is a comma-separated list of NAME=VALUE environment settings for one run (empty string: defaults)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]); reps = int(sys.argv[2])
code = r"""
import ctypes, sys
sys.path.insert(0, %r)
from spwombling_b200 import _lib
lib = _lib.load()
ms = ctypes.c_double(0); info = ctypes.c_int(0)
rc = lib.spw_debug_potrf_time(%d, 1, %d, ctypes.byref(ms), ctypes.byref(info))
print("rc=%%d status=%%d  %%.3f ms per factorisation  (%%.1f TFLOP/s)" %% (rc, info.value, ms.value, %d**3/3/ms.value/1e9))
""" % (root, n, reps, n)
for cfg in sys.argv[3:]:
    env = dict(os.environ)
    for kv in cfg.split(","):
        if kv:
            k, v = kv.split("=")
            env[k] = v
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    print("%-50s %s" % (cfg or "(defaults)", (res.stdout.strip() or res.stderr.strip()[-300:])), flush=True)
