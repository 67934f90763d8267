import os, sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
exec(open('/root/repo/tools/qn_bench.py').read().split("bc.opts.constraint_tol = 0.0")[0])
bc.opts.constraint_tol = 0.0; bc.opts.max_iters = 6
for rep in range(2): bc.project_quasi_newton(qin)
L = C.CDLL(os.environ["MMD_B200_LIB"]); buf = (C.c_longlong * 32)(); L.mmd_debug_timing(buf)
t = list(buf)[:6]
print("n", n, "sweep", t[1]-t[0], "solve+reduce", t[2]-t[1], "alpha", t[3]-t[2], "final/or", t[4]-t[3], "iter total(5-0 of next)", t[5]-t[0])
