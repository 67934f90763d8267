#!/usr/bin/env python
"""Kernel-level timing on the GPU box: ms per launch of the three main kernels after a short burn-in."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains
n = int(os.environ.get("NCH", 4096)); burn = int(os.environ.get("BURN", 20)); dt = float(os.environ.get("DT", 0.1))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy"))
T, S, R = 100, 25, 5
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
bc.opts.solver = int(os.environ.get("SOLVER", 0))   # 0 quasi-Newton, 1 Newton
rng = np.random.default_rng([20200710, 0])
u = rng.standard_normal((n, 4)); v0 = rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
for it in range(burn): bc.hmc_transition(0.05, 8, 1, it)
bc.profile_enable(True, 4096); bc.successful_steps(reset=True)
bc.timer_start()
nst = 16
L = int(os.environ.get("TRAJ", 8))
for tr in range(nst // L):
    bc.transition_begin(1, 1000 + tr)
    bc.transition_steps(dt, L)
    bc.transition_end(1, 1000 + tr, True)
ms = bc.timer_stop_ms()
info = bc.step_info()
res = {"tag": os.environ.get("TAG", ""), "ms_per_step": ms / nst, "chain_steps_per_s": bc.successful_steps() / (ms * 1e-3)}
for kid, nm in [(0, "k_point"), (1, "k_project"), (2, "k_qn"), (3, "k_leapfrog")]:
    c, t = bc.profile_summary(kid); res[nm] = round(t / max(c, 1), 4)
res["iters_fwd_mean"] = float(info["iters_fwd"].mean()); res["iters_fwd_max"] = int(info["iters_fwd"].max())
res["fail"] = float((info["status"] != 0).mean())
print(json.dumps(res))
