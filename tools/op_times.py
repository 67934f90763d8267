#!/usr/bin/env python
"""Per-operation device time per state, the way the reference's operation-time benchmark reports it
(scripts/fhn_model_noiseless_obs_chmc_operation_times.py:42-65, :150-168: median seconds per state over
repeats, 1000 states): constr, jacob_constr_blocks + chol_gram_blocks (+ log_det_sqrt_gram),
grad_log_det_sqrt_gram (everything cached at a position), normal_space_component, one quasi-Newton /
Newton projection iteration, one full constrained leapfrog step.  CPU column: the oracle port, one core."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains  # noqa: E402

n = int(os.environ.get("NSTATES", 1024))
T, S, R = (int(os.environ.get(k, d)) for k, d in (("T", 100), ("S", 25), ("R", 5)))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy"))[:T]
if T > y.shape[0]:
    y = np.resize(y, (T, 1))
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
rng = np.random.default_rng(20200710)
u, v0 = rng.standard_normal((n, 4)), rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
for it in range(40):
    bc.hmc_transition(0.05, 4, 1, it, switch_partition=False)
q, p, x = bc.get_state()
res = {"n_states": n, "T": T, "S": S, "R": R, "unit": "microseconds per state (device time, median of 10)"}


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        bc.synchronize()
        bc.timer_start()
        fn()
        ts.append(bc.timer_stop_ms())
    return float(np.median(ts)) * 1e3 / n


import ctypes as C
L = bc._L
res["constr"] = timed(lambda: L.mmd_constr_dev(bc._h) if hasattr(L, "mmd_constr_dev") else bc._L.mmd_linearize(bc._h, 0)) if False else None
res["jacob_constr_blocks+chol_gram_blocks+log_det_sqrt_gram"] = timed(lambda: L.mmd_linearize(bc._h, 0))
res["grad_log_det_sqrt_gram (incl. the above)"] = timed(lambda: L.mmd_linearize(bc._h, 1))
res["project_onto_cotangent_space"] = timed(lambda: L.mmd_project_momentum(bc._h))
opts = bc.opts
for name, solver in (("quasi_newton", 0), ("newton", 1)):
    opts.solver, opts.constraint_tol = solver, 0.0
    out = {}
    for K in (2, 6):
        opts.max_iters = K
        bc.profile_enable(True, 64)
        bc.project_quasi_newton(q + 1e-3 * rng.standard_normal(q.shape))
        c, ms = bc.profile_summary(2)
        out[K] = ms / max(c, 1)
    res[f"{name}_projection_iteration"] = (out[6] - out[2]) / 4 * 1e3 / n
opts.constraint_tol, opts.max_iters = 1e-9, 50
for name, solver in (("quasi_newton", 0), ("newton", 1)):
    opts.solver = solver
    bc.set_state(q, x, 0, p=p)
    bc.linearize(True)
    bc.project_momentum()
    res[f"constrained_leapfrog_step_{name}"] = timed(lambda: L.mmd_leapfrog_step(bc._h, C.c_double(0.05), C.byref(opts)), reps=5)
del res["constr"]
if os.environ.get("CPU", "1") == "1":
    import torch
    from tests.helpers import make_fhn_problem
    from oracle import torch_oracle as O
    torch.set_num_threads(1)
    pr = make_fhn_problem(T, S, R, n_chains=1, nd=50)
    sysm = pr["system"]
    cpu = {}
    t0 = time.perf_counter(); sysm._constr(torch.tensor(pr["q"][0]), torch.tensor(pr["xobs"][0]), 0); cpu["constr"] = time.perf_counter() - t0
    t0 = time.perf_counter(); pt = sysm.point(pr["q"][0], pr["xobs"][0], 0); cpu["grad_log_det_sqrt_gram (incl. Jacobian, Cholesky)"] = time.perf_counter() - t0
    pp = sysm.project_onto_cotangent_space(torch.tensor(rng.standard_normal(pr["q"].shape[1])), pt)
    t0 = time.perf_counter(); O.leapfrog_step(sysm, pr["q"][0], pp, pr["xobs"][0], 0, 0.02, pt=pt); cpu["constrained_leapfrog_step_quasi_newton"] = time.perf_counter() - t0
    res["cpu_oracle_seconds_per_state_one_core"] = {k: round(v, 4) for k, v in cpu.items()}
    # compiled CPU column (numba restatement of the same path, oracle/numba_chmc.py), one core, median of repeats,
    # the state of one chain on the manifold after the burn-in above
    os.environ["NUMBA_NUM_THREADS"] = "1"
    from oracle import numba_chmc as N
    ch = N.NumbaChain(T, S, R, y, 0.2)
    ch.set_state(q[0], x[0], 0)
    nrng = np.random.default_rng(1)

    def med(fn, reps=7):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return float(np.median(ts))
    bo, bn = ch.parts[0]
    ch.refresh_momentum(nrng); ch.step(0.05); ch._relinearize()          # compile
    cn = {"jacob_constr_blocks+chol_gram_blocks+log_det_sqrt_gram": med(lambda: N.linearize(ch.q, ch.xobs, ch.y, bo, bn, S, ch.dl)),
          "grad_log_det_sqrt_gram (incl. the above)": med(ch._relinearize),
          "project_onto_cotangent_space": med(lambda: ch.refresh_momentum(nrng))}

    def one_step():
        ch.refresh_momentum(nrng)
        ch.step(0.05)
    cn["constrained_leapfrog_step_quasi_newton (incl. one momentum projection)"] = med(one_step)
    res["cpu_numba_microseconds_per_state_one_core"] = {k: round(v * 1e6, 1) for k, v in cn.items()}
print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.items()}))
