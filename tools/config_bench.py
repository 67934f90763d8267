#!/usr/bin/env python
"""Throughput of the other configurations BASELINE.json names (they are parity-test cases, not bench.py lines):
  * FHN noisy observations (noise scale 0.1): T=100, S=40, R=5 (fhn_model_noisy_obs_chmc_experiment.py)
  * SIR, boarding-school data: T=14, S=20, one block of 14 observations, noise scale 1 (sir_model_chmc_experiment.py)
Observation series: the reference's own bundled data sets (scripts/fhn_model_noisy_obs_data.npz,
scripts/sir_model_boarding_school_data.npz) as stored in tests/golden/bundled_configs_golden.npz by
tests/golden/make_golden_bundled.py; chains start from prior draws (FHN) / replicas of the golden initial state with
fresh momenta (SIR).  Static-trajectory transitions (8 leapfrog steps, momentum refresh, accept, partition
switch) timed on the device after a short burn-in; successful chain leapfrog steps per second."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains  # noqa: E402


def run(bc, dt, burn_dt, burn, L=8, ntr=4, tag=""):
    for it in range(burn):
        bc.hmc_transition(burn_dt, L, 1, it)
    bc.successful_steps(reset=True)
    bc.total_qn_iterations(reset=True)
    bc.timer_start()
    for tr in range(ntr):
        bc.hmc_transition(dt, L, 1, 1000 + tr)
    ms = bc.timer_stop_ms()
    ok = bc.successful_steps()
    st = bc.transition_stats()
    return {"config": tag, "chains": bc.n_chains, "dim_q": bc.dim_q, "chains_per_cta_tile": bc.chains_per_tile(),
            "solver": "newton" if bc.opts.solver else "quasi_newton", "step_size": dt,
            "chain_steps_per_s": ok / (ms * 1e-3), "ms_per_leapfrog_step": ms / (L * ntr),
            "step_success_frac": ok / (bc.n_chains * L * ntr), "accept_stat": float(st["accept_stat"].mean()),
            "solver_iterations_per_step": bc.total_qn_iterations() / max(ok, 1)}


def fhn_noisy(n, solver):
    T, S, R = 100, 40, 5
    rng = np.random.default_rng(7)
    g = np.load(os.path.join(ROOT, "tests/golden/bundled_configs_golden.npz"))
    y = np.asarray(g["fhn_noisy_y"], dtype=np.float64)
    assert (int(g["fhn_noisy_T"]), int(g["fhn_noisy_S"]), int(g["fhn_noisy_R"])) == (T, S, R)
    bc = BatchedChains("fhn", float(g["fhn_noisy_obs_interval"]), S, R, y, 4, n, noise=1,
                       sigma_fixed=float(g["fhn_noisy_sigma"]))
    bc.opts.solver = solver
    u = rng.standard_normal((n, 4))
    v0 = rng.standard_normal((n, 2))
    xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
    bc.init_linear_interpolation(u, v0, xo, 0)
    out = run(bc, 0.1, 0.05, 30, tag="FHN noisy obs (bundled data), observation noise 0.1, T=100 S=40 R=5")
    bc.close()
    return out


def sir(n, solver):
    g = np.load(os.path.join(ROOT, "tests/golden/bundled_configs_golden.npz"))
    q = np.tile(g["sir_q0"][None], (n, 1))
    x = np.tile(g["sir_xobs"][None], (n, 1, 1))
    bc = BatchedChains("sir", float(g["sir_obs_interval"]), int(g["sir_S"]), int(g["sir_T"]), g["sir_y"], 4, n, noise=1,
                       sigma_fixed=float(g["sir_sigma"]))
    bc.opts.solver = solver
    bc.set_state(q, x, 0)
    out = run(bc, float(g["sir_dt"]), float(g["sir_dt"]), 20,
              tag="SIR (bundled boarding-school data) T=14 S=20 one block of 14 observations")
    bc.close()
    return out


if __name__ == "__main__":
    n_fhn = int(os.environ.get("NCH_FHN", 8192))
    n_sir = int(os.environ.get("NCH_SIR", 131072))
    for solver in (0, 1):
        print(json.dumps(fhn_noisy(n_fhn, solver)), flush=True)
        print(json.dumps(sir(n_sir, solver)), flush=True)
