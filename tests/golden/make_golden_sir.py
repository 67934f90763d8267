#!/usr/bin/env python
"""Golden vectors for the SIR experiment shape of the reference (scripts/sir_model_chmc_experiment.py:
14 daily observations, 20 steps per observation, one block of all observations, inferred observation
noise scale, dim_q = 860) from the float64 autodiff oracle, both projection solvers.  ORACLE-frozen:
they pin the CUDA path to the restatement (the reference itself cannot run here).

    python tests/golden/make_golden_sir.py      # ~2 minutes on one core; writes sir_T14_S20_golden.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as O  # noqa: E402
from oracle.models import sir  # noqa: E402

T, S, N_CHAINS, DT = 14, 20, 2, 0.02
# a synthetic 14-day epidemic curve of the boarding-school data's shape (NOT the bundled series 3, 8, 28, 75, 221,
# 281, ...; that one is covered by make_golden_bundled.py / test_gpu_bundled_configs.py)
y = np.array([3, 8, 26, 76, 225, 298, 258, 233, 189, 128, 68, 29, 14, 4], dtype=np.float64)[:, None]
sysm = O.OracleSystem(1.0, S, T, y, 5, 3, 3, sir.forward_func, sir.generate_x_0, sir.generate_z, sir.obs_func,
                      sir.generate_σ_y, False, dim_v_0=1)
out = {"T": T, "S": S, "dt": DT, "y": y}
q0s, xos, praws, lds, grads, nscs, cs = [], [], [], [], [], [], []
traj = {"quasi_newton": [], "newton": []}
for c in range(N_CHAINS):
    rng = np.random.default_rng([20200710, 100 + c])
    u = np.array([-1.0, -0.5, 0.8, 0.0, np.log(5.0)]) + 0.1 * rng.standard_normal(5)
    v0 = np.array([0.8 + 0.1 * rng.standard_normal()])
    v = 0.2 * rng.standard_normal((T * S, 3))
    z = sir.generate_z(torch.tensor(u[:4]))
    sig = float(np.exp(u[4]))
    x = sir.generate_x_0(z, torch.tensor(v0))
    xobs, n = [], []
    for t in range(T * S):
        x = sir.forward_func(z, x, torch.tensor(v[t]), 1.0 / S)
        if (t + 1) % S == 0:
            xobs.append(x.numpy().copy())
            n.append((y[(t + 1) // S - 1, 0] - float(torch.exp(x[1]))) / sig)   # on the manifold
    q = np.concatenate([u, v0, v.reshape(-1), np.array(n)])
    xo = np.stack(xobs)
    assert xo[:, :2].min() > -20
    p_raw = rng.standard_normal(q.shape[0])
    pt = sysm.point(q, xo, 0)
    q0s.append(q); xos.append(xo); praws.append(p_raw)
    cs.append(sysm._constr(torch.tensor(q), torch.tensor(xo), 0).numpy())
    lds.append(pt["ld"]); grads.append(pt["grad_ld"].numpy())
    nscs.append(sysm._normal_space_component(torch.tensor(p_raw), pt["jac"], pt["chol"]).numpy())
    for solver in traj:
        p = sysm.project_onto_cotangent_space(torch.tensor(p_raw), pt)
        qn, pn, ptn, inf = O.leapfrog_step(sysm, q, p, xo, 0, DT, pt=pt, solver=solver)
        traj[solver].append((qn.numpy(), pn.numpy(), sysm.h(qn, pn, ptn), inf["n_fwd"], inf["n_back"]))
    print("chain", c, "ld", pt["ld"], "iters", [(t[-1][3], t[-1][4]) for t in traj.values()], flush=True)
out.update(q0=np.stack(q0s), xobs=np.stack(xos), p_raw=np.stack(praws), c=np.stack(cs), ld=np.array(lds),
           grad_ld=np.stack(grads), nsc=np.stack(nscs))
for solver, tr in traj.items():
    out[f"{solver}_q"] = np.stack([t[0] for t in tr])
    out[f"{solver}_p"] = np.stack([t[1] for t in tr])
    out[f"{solver}_h"] = np.array([t[2] for t in tr])
    out[f"{solver}_it"] = np.array([[t[3], t[4]] for t in tr])
np.savez_compressed(os.path.join(HERE, "sir_T14_S20_golden.npz"), **out)
print("wrote sir_T14_S20_golden.npz")
