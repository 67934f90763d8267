#!/usr/bin/env python
"""Golden vectors on the reference's BUNDLED data sets, at the defaults of its experiment scripts:

  config 2  scripts/fhn_model_noisy_obs_chmc_experiment.py:61-64 -- fhn_model_noisy_obs_simulated_data.npz,
            y = y_seq_mean + 0.1 n_seq, T=100, S=40, R=5, fixed observation noise 0.1, Newton projection;
            chain started by find_initial_state_by_linear_interpolation (:104-118)
  config 3  scripts/sir_model_chmc_experiment.py:62-79 -- sir_model_boarding_school_data.npz (14 daily counts),
            S=20, one block of all 14 observations, fixed observation noise 1.0, Newton projection

from the float64 autodiff oracle (constraint, log-det, its gradient, normal-space component, one constrained
leapfrog step with both solvers).  The .npz files are read from the reference checkout (read-only, only here); the
observations are stored in the fixture so the GPU tests need nothing outside the repo.  ORACLE-frozen.

    python tests/golden/make_golden_bundled.py     # a few minutes on one core; writes bundled_configs_golden.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as O  # noqa: E402
from oracle.models import fhn, sir  # noqa: E402

REF = "/root/reference/scripts"
out = {}


def record(tag, sysm, q, xo, dt, seed):
    rng = np.random.default_rng(seed)
    p_raw = rng.standard_normal(q.shape[0])
    pt = sysm.point(q, xo, 0)
    out[f"{tag}_q0"], out[f"{tag}_xobs"], out[f"{tag}_p_raw"] = q, xo, p_raw
    out[f"{tag}_c"] = sysm._constr(torch.tensor(q), torch.tensor(xo), 0).numpy()
    out[f"{tag}_ld"] = float(pt["ld"])
    out[f"{tag}_grad_ld"] = pt["grad_ld"].numpy()
    out[f"{tag}_nsc"] = sysm._normal_space_component(torch.tensor(p_raw), pt["jac"], pt["chol"]).numpy()
    out[f"{tag}_dt"] = dt
    for solver in ("newton", "quasi_newton"):
        p = sysm.project_onto_cotangent_space(torch.tensor(p_raw), pt)
        qn, pn, ptn, inf = O.leapfrog_step(sysm, q, p, xo, 0, dt, pt=pt, solver=solver)
        out[f"{tag}_{solver}_q"], out[f"{tag}_{solver}_p"] = qn.numpy(), pn.numpy()
        out[f"{tag}_{solver}_h"] = float(sysm.h(qn, pn, ptn))
        out[f"{tag}_{solver}_it"] = np.array([inf["n_fwd"], inf["n_back"]])
        print(tag, solver, "iters", inf["n_fwd"], inf["n_back"], "h", out[f"{tag}_{solver}_h"], flush=True)


# ---- config 2: FHN, noisy observations, bundled simulated data ----
d = np.load(os.path.join(REF, "fhn_model_noisy_obs_simulated_data.npz"))
sigma = 0.1
y = (d["y_seq_mean"] + sigma * d["n_seq"])[:, None]
T, S, R = int(d["num_obs"]), 40, 5
sysm = O.OracleSystem(float(d["obs_interval"]), S, R, y, 4, 2, 2, fhn.forward_func, fhn.generate_x_0, fhn.generate_z,
                      fhn.obs_func, sigma, False, dim_v_0=2)
rng = np.random.default_rng(20200710)
q, xo = O.find_initial_state_by_linear_interpolation(
    sysm, rng, lambda r: np.concatenate((y, r.standard_normal(y.shape) * 0.5), -1),
    u=0.5 * rng.standard_normal(4), v_0=rng.standard_normal(2))
out.update(fhn_noisy_y=y, fhn_noisy_T=T, fhn_noisy_S=S, fhn_noisy_R=R, fhn_noisy_sigma=sigma,
           fhn_noisy_obs_interval=float(d["obs_interval"]))
record("fhn_noisy", sysm, q.numpy(), xo.numpy(), 0.02, [1, 2])

# ---- config 3: SIR, boarding-school data ----
d = np.load(os.path.join(REF, "sir_model_boarding_school_data.npz"))
y = np.asarray(d["y_seq"], dtype=np.float64)
T, S, sigma = int(d["num_obs"]), 20, 1.0
sysm = O.OracleSystem(float(d["obs_interval"]), S, T, y, 4, 3, 3, sir.forward_func, sir.generate_x_0, sir.generate_z,
                      sir.obs_func, sigma, False, dim_v_0=1)
rng = np.random.default_rng([20200710, 3])
u = np.array([-1.0, -0.5, 0.8, 0.0]) + 0.1 * rng.standard_normal(4)
v0 = np.array([0.8 + 0.1 * rng.standard_normal()])
v = 0.2 * rng.standard_normal((T * S, 3))
z = sir.generate_z(torch.tensor(u))
x = sir.generate_x_0(z, torch.tensor(v0))
xobs, n = [], []
for t in range(T * S):
    x = sir.forward_func(z, x, torch.tensor(v[t]), float(d["obs_interval"]) / S)
    if (t + 1) % S == 0:
        xobs.append(x.numpy().copy())
        n.append((y[(t + 1) // S - 1, 0] - float(torch.exp(x[1]))) / sigma)    # noise that puts the state on the manifold
q = np.concatenate([u, v0, v.reshape(-1), np.array(n)])
out.update(sir_y=y, sir_T=T, sir_S=S, sir_sigma=sigma, sir_obs_interval=float(d["obs_interval"]))
record("sir", sysm, q, np.stack(xobs), 0.01, [3, 4])

np.savez_compressed(os.path.join(HERE, "bundled_configs_golden.npz"), **out)
print("wrote bundled_configs_golden.npz")
