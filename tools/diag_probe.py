import ctypes, os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import spwombling_b200 as spw
from spwombling_b200 import _lib
import oracle
lib = _lib.load()
buf = (ctypes.c_longlong * 16)()
lib.spw_debug_diag_timing(buf)
n = 900
rng = np.random.Generator(np.random.Philox(1)); coords = rng.random((n, 2))
A = oracle.cov_matern1(oracle.dist_matrix(coords), 8.0, 2.0) + 0.5 * np.eye(n)
for _ in range(3): spw.chol(A)
lib.spw_debug_diag_timing(buf)
print("cycles: load %d, sweep end %d, stores end %d  -> sweep %.2f us at 1.965 GHz" % (buf[0], buf[1], buf[4], (buf[1]-buf[0])/1965.0))
