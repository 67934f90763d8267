#!/usr/bin/env python
"""Distribution of quasi-Newton iteration counts / failure kinds in the burned-in regime."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains
n = int(os.environ.get("NCH", 4096)); burn = int(os.environ.get("BURN", 150)); dt = float(os.environ.get("DT", 0.1))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy"))
T, S, R = 100, 25, 5
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
rng = np.random.default_rng([20200710, 0])
u = rng.standard_normal((n, 4)); v0 = rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
for it in range(burn): bc.hmc_transition(0.05, 8, 1, it)
hf = np.zeros(52, int); hr = np.zeros(52, int); stc = {}
acc = []
for it in range(12):
    bc.transition_begin(1, 1000 + it)
    for s in range(8):
        live_before = bc.step_info()["status"] == 0 if s else np.ones(n, bool)
        bc.transition_steps(dt, 1)
        info = bc.step_info()
        f = info["iters_fwd"][live_before]; hf += np.bincount(f, minlength=52)[:52]
        ok_fwd = live_before & ((info["status"] & 3) == 0)
        r = info["iters_rev"][ok_fwd]; hr += np.bincount(r, minlength=52)[:52]
        for v in info["status"][live_before]: stc[int(v)] = stc.get(int(v), 0) + 1
    bc.transition_end(1, 1000 + it, True)
    acc.append(bc.transition_stats()["accept_stat"].mean())
q, _, _ = bc.get_state()
print(json.dumps({"dt": dt, "hist_fwd": hf.tolist(), "hist_rev": hr.tolist(), "status_counts": stc, "accept_stat": float(np.mean(acc)),
                  "half_q2": float(0.5 * (q ** 2).sum(1).mean()), "z_mean": np.exp(q[:, :3]).mean(0).round(3).tolist()}))
