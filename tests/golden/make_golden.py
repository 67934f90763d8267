#!/usr/bin/env python
"""Generate the canonical-size golden vectors (FHN noiseless, T=100, S=25, R=5) from the float64
autodiff oracle.  The reference itself cannot run in this environment (mici / jax / symnum are not
installable), so these are ORACLE-frozen vectors; since round 2 the file is cross-checked against the reference's own
source executed through torch / SymPy stand-ins (tests/test_reference_pin_cpu.py::
test_canonical_size_golden_equals_the_reference).

    python tests/golden/make_golden.py      # ~3 minutes on one core; writes fhn_T100_S25_R5_golden.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as O  # noqa: E402
from oracle.models import fhn  # noqa: E402

T, S, R = 100, 25, 5
N_CHAINS, N_STEPS, DT = 2, 2, 0.05
y = np.load(os.path.join(HERE, "fhn_yseq_T100.npy"))
sysm = O.OracleSystem(0.2, S, R, y, 4, 2, 2, fhn.forward_func, fhn.generate_x_0, fhn.generate_z, fhn.obs_func,
                      None, False, dim_v_0=2)


def gen_init(rng):
    return np.concatenate((y, rng.standard_normal(y.shape) * 0.5), -1)


out = {"T": T, "S": S, "R": R, "dt": DT, "y": y}
q0s, xos, praws, couts, lds, grads, nscs = [], [], [], [], [], [], []
traj_q, traj_p, traj_h, traj_it = [], [], [], []
for c in range(N_CHAINS):
    rng = np.random.default_rng([20200710, c])
    q, xo = O.find_initial_state_by_linear_interpolation(sysm, rng, gen_init, u=0.5 * rng.standard_normal(4),
                                                         v_0=rng.standard_normal(2))
    part = c % 2
    p_raw = torch.tensor(rng.standard_normal(q.shape[0]))
    q_off = q + 0.01 * torch.tensor(rng.standard_normal(q.shape[0]))
    couts.append(sysm._constr(q_off, xo, part).numpy())
    pt = sysm.point(q, xo, part)
    lds.append(pt["ld"])
    grads.append(pt["grad_ld"].numpy())
    nscs.append(sysm._normal_space_component(p_raw, pt["jac"], pt["chol"]).numpy())
    p = sysm.project_onto_cotangent_space(p_raw, pt)
    qs, ps, hs, its = [], [], [sysm.h(q, p, pt)], []
    qq = q
    for s in range(N_STEPS):
        qq, p, pt, info = O.leapfrog_step(sysm, qq, p, xo, part, DT, pt=pt)
        qs.append(qq.numpy()); ps.append(p.numpy()); hs.append(sysm.h(qq, p, pt))
        its.append([info["n_fwd"], info["n_back"]])
        print("chain", c, "step", s, info, flush=True)
    q0s.append(q.numpy()); xos.append(xo.numpy()); praws.append(p_raw.numpy())
    traj_q.append(np.stack(qs)); traj_p.append(np.stack(ps)); traj_h.append(np.array(hs)); traj_it.append(np.array(its))
out.update(q0=np.stack(q0s), xobs=np.stack(xos), p_raw=np.stack(praws), c_off=np.array(couts, dtype=object),
           q_off_noise_seed=0, ld=np.array(lds), grad_ld=np.stack(grads), nsc=np.stack(nscs),
           traj_q=np.stack(traj_q), traj_p=np.stack(traj_p), traj_h=np.stack(traj_h), traj_it=np.stack(traj_it))
# c_off has different lengths per partition (119 / 120): store separately
del out["c_off"]
for c in range(N_CHAINS):
    out[f"c_off_{c}"] = couts[c]
np.savez_compressed(os.path.join(HERE, "fhn_T100_S25_R5_golden.npz"), **out)
print("wrote golden")
