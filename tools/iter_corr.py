#!/usr/bin/env python
"""Are per-chain projection iteration counts persistent (so that sorting chains into tiles by their
last count would shrink the lockstep tail)?  Prints correlations between consecutive steps / transitions
and the lockstep efficiency mean / E[max over tile] for natural vs sorted tiling."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains
n = int(os.environ.get("NCH", 4096)); dt = float(os.environ.get("DT", 0.1))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy")); T, S, R = 100, 25, 5
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
rng = np.random.default_rng([20200710, 0])
u = rng.standard_normal((n, 4)); v0 = rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
for it in range(20): bc.hmc_transition(0.05, 8, 1, it)
its = []
for tr in range(6):
    bc.transition_begin(1, 1000 + tr)
    for s in range(4):
        bc.transition_steps(dt, 1)
        info = bc.step_info()
        its.append(np.where(info["status"] == 0, info["iters_fwd"] + info["iters_rev"], 100).astype(float))
    bc.transition_end(1, 1000 + tr, True)
its = np.array(its)  # [24, n]
def eff(x, order, cpb=8):
    xs = x[order].reshape(-1, cpb)
    return float(x.mean() / xs.max(1).mean())
res = {"mean_iters": float(its.mean()), "corr_step": float(np.corrcoef(its[0], its[1])[0, 1]),
       "corr_next_transition": float(np.corrcoef(its[3], its[4])[0, 1]), "corr_far": float(np.corrcoef(its[0], its[-1])[0, 1])}
nat = np.arange(n)
res["eff_natural_cpb8"] = float(np.mean([eff(its[i], nat) for i in range(1, 24)]))
res["eff_sorted_by_prev_cpb8"] = float(np.mean([eff(its[i], np.argsort(its[i - 1], kind="stable")) for i in range(1, 24)]))
res["eff_sorted_by_prev_cpb32"] = float(np.mean([eff(its[i], np.argsort(its[i - 1], kind="stable"), 32) for i in range(1, 24)]))
res["eff_natural_cpb32"] = float(np.mean([eff(its[i], nat, 32) for i in range(1, 24)]))
res["eff_natural_cpb4"] = float(np.mean([eff(its[i], nat, 4) for i in range(1, 24)]))
res["eff_oracle_sorted_cpb8"] = float(np.mean([eff(its[i], np.argsort(its[i])) for i in range(1, 24)]))
print(json.dumps(res))
