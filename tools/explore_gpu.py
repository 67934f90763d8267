#!/usr/bin/env python
"""Exploration on the GPU box: burn-in behaviour, step-size vs acceptance, time per step."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains

n = int(os.environ.get("NCH", 4096))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy"))
T, S, R = 100, 25, 5
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
rng = np.random.default_rng(20200710)
u = rng.standard_normal((n, 4)); v0 = rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
c = bc.constr(); print("init max|c|", np.abs(c).max())
bc.linearize(True)
bc.sample_momentum(1, 0)
print("H0 mean", bc.hamiltonian().mean())
L = 8
for dt in [0.02, 0.05, 0.1]:
    pass
dt = float(os.environ.get("DT", 0.05))
t0 = time.time()
for it in range(int(os.environ.get("NIT", 60))):
    bc.timer_start()
    bc.hmc_transition(dt, L, 1234, it)
    ms = bc.timer_stop_ms()
    st = bc.transition_stats(); info = bc.step_info()
    if it % 5 == 0 or it > 54:
        bc.linearize(False)
        q, _, _ = bc.get_state()
        print(it, "ms/transition %.1f" % ms, "acc %.3f" % st["accepted"].mean(), "accstat %.3f" % st["accept_stat"].mean(),
              "fail %.3f" % (st["status"] != 0).mean(), "it_fwd %.2f it_rev %.2f" % (info["iters_fwd"].mean(), info["iters_rev"].mean()),
              "0.5|q|^2 %.0f" % (0.5 * (q ** 2).sum(1).mean()), "z0 %s" % np.exp(q[:, :3]).mean(0).round(3), flush=True)
print("wall", time.time() - t0)
# step-size scan from the burned-in state
for dt in [0.05, 0.1, 0.15, 0.2, 0.3]:
    accs, fails, itf, mss = [], [], [], []
    for it in range(6):
        bc.timer_start()
        bc.hmc_transition(dt, L, 999, 1000 + it)
        mss.append(bc.timer_stop_ms())
        st = bc.transition_stats(); info = bc.step_info()
        accs.append(st["accept_stat"].mean()); fails.append((st["status"] != 0).mean()); itf.append(info["iters_fwd"].mean())
    print("dt", dt, "accept_stat %.3f fail %.3f iters_fwd(last step) %.2f ms/transition %.1f" % (np.mean(accs), np.mean(fails), np.mean(itf), np.mean(mss)), flush=True)
