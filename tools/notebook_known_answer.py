#!/usr/bin/env python
"""Distributional known-answer test against the reference's own recorded output.

FitzHugh-Nagumo_example.ipynb fully specifies a posterior: latent vector q_ref =
RandomState(20200710).standard_normal(5006) pushed through the model (cells 18-27: notebook prior
parametrisation, strong-order-1.5 step, obs_interval 0.5, 100 observations x 25 steps) gives y_seq; the
notebook then samples the conditioned diffusion with Gaussian splitting, the Newton projection solver
(tol 1e-9 / 1e-8, reverse check 2e-8) and partition switching, and records the ArviZ table of cell 45.
Any correct sampler of the same posterior must reproduce those means within Monte Carlo error.  Here
many chains run the on-device static-trajectory constrained HMC transition (momentum refresh, L
constrained leapfrog steps, Metropolis accept, partition switch) of libmmd_b200.so on that problem."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains  # noqa: E402
from manifold_mcmc_for_diffusions_b200.diagnostics import ess_bulk, rhat as split_rhat  # noqa: E402
from manifold_mcmc_for_diffusions_b200.example_models import fhn_notebook as m  # noqa: E402

# ArviZ summary recorded in the notebook (cell 45): mean, sd, mcse_mean
NOTEBOOK = {"σ": (0.784, 0.073, 0.003), "ϵ": (0.206, 0.028, 0.001), "γ": (1.054, 0.149, 0.004),
            "β": (0.282, 0.110, 0.002), "x_0[0]": (-0.202, 0.970, 0.023), "x_0[1]": (1.289, 0.574, 0.011)}


def main():
    n = int(os.environ.get("NCH", 512))
    n_burn = int(os.environ.get("NBURN", 400))
    n_main = int(os.environ.get("NMAIN", 400))
    L = int(os.environ.get("L", 10))
    dt = float(os.environ.get("DT", 0.15))
    solver = int(os.environ.get("SOLVER", 1))
    T, S, R, obs_interval = 100, 25, 5, 0.5
    q_ref = np.random.RandomState(20200710).standard_normal(m.dim_z + m.dim_x + T * S * m.dim_v)
    _, y_seq, z_ref, x0_ref = m.generate_from_model(q_ref, obs_interval / S, S)
    bc = BatchedChains("fhn_notebook", obs_interval, S, R, y_seq, 4, n, use_gaussian_splitting=True)
    bc.opts.solver = solver
    bc.opts.constraint_tol, bc.opts.position_tol, bc.opts.reverse_check_tol = 1e-9, 1e-8, 2e-8
    rng = np.random.default_rng(20200710)
    # initial parameters: prior draws shrunk towards the prior mean (any over-dispersed start is valid for MCMC;
    # unshrunk prior draws put ~30 % of the chains at epsilon < 0.08 where delta / epsilon makes the discretised
    # dynamics so stiff that every trajectory fails at any step size and a fixed-step-size chain never leaves)
    u_scale = float(os.environ.get("USCALE", 0.4))
    u, v0 = u_scale * rng.standard_normal((n, 4)), rng.standard_normal((n, 2))
    xo = np.concatenate((np.broadcast_to(y_seq, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
    bc.init_linear_interpolation(u, v0, xo, 0)
    t0 = time.time()
    it = 0
    log = []
    # staged burn-in: the interpolated initial states are far in the tails.  A chain whose initial state lies where
    # the constrained integrator fails at every step size (every trajectory rejected) never moves with a
    # fixed step size shared by all chains; after each stage such chains are restarted from the state of a
    # randomly chosen moving chain (a valid re-initialisation: burn-in continues afterwards and the per-chain
    # Philox momentum streams decorrelate the copies at once).
    n_restarted = 0
    adapt = int(os.environ.get("ADAPT", 0))
    if adapt:
        # per-chain dual averaging on device (target 0.8, reg 0.1 as in scripts/utils.py:303-306), pooled at the end
        bc.adapt_start(dt / 4, target=0.8, reg_coefficient=0.1)
    for frac, step in ((0.25, dt / 4), (0.25, dt / 2), (0.5, dt)):
        acc, moved = [], np.zeros(n)
        for _ in range(int(frac * n_burn)):
            bc.hmc_transition(step, L, 20200710, it)
            it += 1
            st = bc.transition_stats()
            acc.append(st["accept_stat"].mean())
            moved += st["accepted"]
        stuck = np.flatnonzero(moved == 0)
        if len(stuck) and len(stuck) < n:
            q, _, x = bc.get_state()
            alive = np.flatnonzero(moved > 0)
            src = rng.choice(alive, size=len(stuck))
            q[stuck], x[stuck] = q[src], x[src]
            bc.set_state(q, x, bc.partition)
            n_restarted += len(stuck)
        log.append({"dt": step, "transitions": len(acc), "accept_stat": float(np.mean(acc[-20:])),
                    "restarted_chains": int(len(stuck))})
    if adapt:
        per_chain = bc.get_step_sizes()
        bc.adapt_stop(pool=True)
        log.append({"adapted_step_size_quantiles": [round(float(v), 4) for v in np.quantile(per_chain, [0.05, 0.5, 0.95])],
                    "pooled_step_size": float(bc.get_step_sizes()[0])})
    names = ["σ", "ϵ", "γ", "β", "x_0[0]", "x_0[1]"]
    draws = np.empty((n, n_main, 6))
    acc, fail = [], []
    acc_chain = np.zeros(n)
    bits = np.zeros(4)
    t_main = time.time()
    for k in range(n_main):
        bc.hmc_transition(dt, L, 20200710, it)
        it += 1
        st = bc.transition_stats()
        acc.append(st["accept_stat"].mean())
        acc_chain += st["accepted"]
        fail.append((st["status"] != 0).mean())
        for b in range(4):
            bits[b] += ((st["status"] >> b) & 1).mean()
        q, _, _ = bc.get_state()
        z = m.generate_z(q[:, :4])
        draws[:, k, :4] = z
        draws[:, k, 4:] = m.generate_x_0(z, q[:, 4:6])
    wall = time.time() - t0
    main_wall = time.time() - t_main
    out = {"chains": n, "burn_in": log, "main_transitions": n_main, "leapfrog_per_transition": L, "dt": dt,
           "solver": "newton" if solver else "quasi_newton", "accept_stat": float(np.mean(acc)),
           "integrator_error_rate": float(np.mean(fail)),
           "error_bits_not_converged_diverged_nonreversible_nonfinite": [round(float(b / n_main), 4) for b in bits], "wall_s": round(wall, 1),
           "truth": dict(zip(names, [float(a) for a in list(z_ref) + list(x0_ref)])), "vars": {}}
    worst = 0.0
    for j, nm in enumerate(names):
        x = draws[:, :, j]
        chain_means = x.mean(1)
        mean, sd = float(x.mean()), float(x.std())
        mcse = float(chain_means.std(ddof=1) / np.sqrt(n))
        ref_mean, ref_sd, ref_mcse = NOTEBOOK[nm]
        zscore = (mean - ref_mean) / np.hypot(ref_mcse, mcse)
        worst = max(worst, abs(zscore))
        out["vars"][nm] = {"mean": round(mean, 4), "sd": round(sd, 4), "mcse": round(mcse, 5),
                           "rhat": round(float(split_rhat(x)), 4), "ess_bulk": round(float(ess_bulk(x)), 0),
                           "notebook_mean": ref_mean, "notebook_sd": ref_sd, "notebook_mcse": ref_mcse,
                           "z": round(float(zscore), 2)}
    out["max_abs_z"] = round(worst, 2)
    # effective samples per second of the main phase (wall clock incl. the per-transition trace read-back):
    # rank-normalised bulk ESS pooled over chains, minimum over the traced variables
    ess_min = min(v["ess_bulk"] for v in out["vars"].values())
    out["ess"] = {"min_bulk_ess": ess_min, "main_phase_wall_s": round(main_wall, 2),
                  "ess_per_s": round(ess_min / main_wall, 1),
                  "ess_per_chain_leapfrog_step": ess_min / (n * n_main * L),
                  "chain_leapfrog_steps_per_s": round(n * n_main * L / main_wall, 0),
                  "notebook": "2 chains x 750 NUTS transitions in 591 s on the authors' CPU (cell 43 output)"}
    out["stuck_chain_fraction"] = float(np.mean(acc_chain == 0))
    out["init_u_scale"] = u_scale
    out["restarted_chains_during_burn_in"] = int(n_restarted)
    if os.environ.get("VERBOSE"):
        qs = [0.05, 0.25, 0.5, 0.75, 0.95]
        out["per_chain_accept_rate_quantiles"] = [round(float(v), 3) for v in np.quantile(acc_chain / n_main, qs)]
        for j, nm in enumerate(names[:2]):
            h1 = draws[:, : n_main // 2, j].mean(1)
            h2 = draws[:, n_main // 2:, j].mean(1)
            out["chain_mean_quantiles_" + nm] = {"first_half": [round(float(v), 3) for v in np.quantile(h1, qs)],
                                                 "second_half": [round(float(v), 3) for v in np.quantile(h2, qs)],
                                                 "within_chain_sd_median": round(float(np.median(draws[:, :, j].std(1))), 4)}
    print(json.dumps(out, ensure_ascii=False))


if __name__ == "__main__":
    main()
