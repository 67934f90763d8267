#!/usr/bin/env python
"""Per-iteration cost of the on-device quasi-Newton loop with every chain forced to run exactly K
iterations (constraint_tol = 0 never passes), so there is no tail: isolates sweep + block solves."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains
n = int(os.environ.get("NCH", 4096))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy"))
T, S, R = 100, 25, 5
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
rng = np.random.default_rng([20200710, 0])
u = rng.standard_normal((n, 4)); v0 = rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
bc.linearize(True)
q, _, _ = bc.get_state()
qin = q + 1e-3 * rng.standard_normal(q.shape)
bc.opts.constraint_tol = 0.0
res = {"tag": os.environ.get("TAG", ""), "n": n}
import ctypes as C
from manifold_mcmc_for_diffusions_b200.batched import _dp, _ip
out = np.empty_like(qin); st = np.empty(n, dtype=np.int32); it = np.empty(n, dtype=np.int32)
for K in [int(k) for k in os.environ.get("KLIST", "1,6,16").split(",")]:
    bc.opts.max_iters = K
    ts = []
    for rep in range(3):
        bc.profile_enable(True, 64)
        bc.project_quasi_newton(qin)
        c, ms = bc.profile_summary(2)
        ts.append(ms / max(c, 1))
    res["K%d_ms" % K] = round(min(ts), 4)
if "K16_ms" in res and "K6_ms" in res: res["ms_per_iter"] = round((res["K16_ms"] - res["K6_ms"]) / 10.0, 4)
if "ms_per_iter" in res: res["GBps_sweep"] = round(n * 2500 * 6 * 8 / (res["ms_per_iter"] * 1e-3) / 1e9, 1)
print(json.dumps(res))
