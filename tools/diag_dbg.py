import sys, ctypes, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, spwombling_b200 as spw
lib = spw.load()
buf = (ctypes.c_longlong*16)()
lib.spw_debug_diag_timing(ctypes.cast(buf, ctypes.c_void_p))
rng=np.random.default_rng(0); n=640
c=rng.random((n,2)); import oracle
A=oracle.cov_matern2(oracle.dist_matrix(c),6.0,1.0)+np.eye(n)
for rep in range(3):
    spw.chol(A)
    lib.spw_debug_diag_timing(ctypes.cast(buf, ctypes.c_void_p))
    v=list(buf); print("cycles: load %d factor+inverse sweep %d write-out %d total %d (%.1f us @1.9GHz)"%(v[0], v[1]-v[0], v[4]-v[1], v[4], v[4]/1.9e3))
