#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN SOURCE (`/root/reference/sde/mici_extensions.py`, executed unmodified
through oracle/reference_runner.py: torch-backed `jax` stand-in, minimal `mici` stand-in, Mici's constrained leapfrog
step order, the reference's projection-solver wrappers) for the GPU parity test tests/test_gpu_reference_pin.py.
Runs only where the reference checkout exists (the build container):

    python tests/golden/make_golden_reference_pin.py      # writes tests/golden/reference_pin_golden.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_runner as R  # noqa: E402
from tests.helpers import make_fhn_problem  # noqa: E402
from manifold_mcmc_for_diffusions_b200.mici_compat.integrators import ConstrainedLeapfrogIntegrator  # noqa: E402

CASES = [dict(tag="noiseless", T=10, S=5, R=5, noise=0, gaussian=False),
         dict(tag="ragged_gauss", T=12, S=4, R=5, noise=0, gaussian=True),
         dict(tag="fixed_noise", T=10, S=5, R=5, noise=1, gaussian=False),
         dict(tag="inferred_noise", T=10, S=5, R=5, noise=2, gaussian=False)]
TOL = dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10, max_iters=50)
DT, N_STEPS, N_CHAINS = 0.05, 2, 3

out = {"dt": DT, "n_steps": N_STEPS, "tags": np.array([c["tag"] for c in CASES])}
ref = R.load()
for c in CASES:
    prob = make_fhn_problem(c["T"], c["S"], c["R"], n_chains=N_CHAINS, nd=200, noise=c["noise"], gaussian=c["gaussian"])
    sysr = R.make_fhn_system(0.2, c["S"], c["R"], prob["y"], noise=c["noise"], sigma=prob["sigma"],
                             use_gaussian_splitting=c["gaussian"])
    t = c["tag"]
    rng = np.random.default_rng(17)
    p_raw = rng.standard_normal(prob["q"].shape)
    out.update({f"{t}_T": c["T"], f"{t}_S": c["S"], f"{t}_R": c["R"], f"{t}_noise": c["noise"],
                f"{t}_gaussian": int(c["gaussian"]), f"{t}_sigma": prob["sigma"], f"{t}_y": prob["y"],
                f"{t}_q0": prob["q"], f"{t}_xobs": prob["xobs"], f"{t}_p_raw": p_raw})
    for part in (0, 1):
        cs, lds, gs, ps, hs = [], [], [], [], []
        for i in range(N_CHAINS):
            st = ref.ConditionedDiffusionHamiltonianState(pos=prob["q"][i].copy(), x_obs_seq=prob["xobs"][i], partition=part)
            cs.append(np.asarray(sysr.constr(st), dtype=np.float64))
            lds.append(float(sysr.log_det_sqrt_gram(st)))
            gs.append(np.asarray(sysr.grad_log_det_sqrt_gram(st), dtype=np.float64))
            st.mom = np.asarray(sysr.project_onto_cotangent_space(p_raw[i].copy(), st), dtype=np.float64)
            ps.append(st.mom.copy())
            hs.append(float(sysr.h(st)))
        out.update({f"{t}_p{part}_c": np.stack(cs), f"{t}_p{part}_ld": np.array(lds), f"{t}_p{part}_grad_ld": np.stack(gs),
                    f"{t}_p{part}_p0": np.stack(ps), f"{t}_p{part}_h0": np.array(hs)})
        for solver, wrapper in (("quasi_newton", ref.jitted_solve_projection_onto_manifold_quasi_newton),
                                ("newton", ref.jitted_solve_projection_onto_manifold_newton)):
            integ = ConstrainedLeapfrogIntegrator(sysr, step_size=DT, n_inner_step=1, reverse_check_tol=2e-8,
                                                  projection_solver=wrapper, projection_solver_kwargs=TOL)
            qs, pps, hh = [], [], []
            for i in range(N_CHAINS):
                st = ref.ConditionedDiffusionHamiltonianState(pos=prob["q"][i].copy(), x_obs_seq=prob["xobs"][i],
                                                              partition=part, mom=ps[i].copy())
                for k_step in range(N_STEPS):
                    st = integ.step(st)
                    if i == 0 and k_step == 0:
                        # Mici's per-method call counts after ONE step from a fresh state (the paper's cost accounting;
                        # the solver wrappers book their iterations on them, :1382-1387, :1451-1461)
                        cc = {k[2]: int(v) for k, v in st._call_counts.items()}
                        out[f"{t}_p{part}_{solver}_count_names"] = np.array(sorted(cc))
                        out[f"{t}_p{part}_{solver}_count_values"] = np.array([cc[k] for k in sorted(cc)])
                qs.append(np.asarray(st.pos, dtype=np.float64)); pps.append(np.asarray(st.mom, dtype=np.float64))
                hh.append(float(sysr.h(st)))
            out.update({f"{t}_p{part}_{solver}_q": np.stack(qs), f"{t}_p{part}_{solver}_p": np.stack(pps),
                        f"{t}_p{part}_{solver}_h": np.array(hh)})
    # SwitchPartitionTransition: x_obs_seq regenerated from the position (:1262-1282)
    st = ref.ConditionedDiffusionHamiltonianState(pos=prob["q"][0].copy(), x_obs_seq=prob["xobs"][0], partition=0)
    ref.SwitchPartitionTransition(sysr).sample(st, None)
    out[f"{t}_switch_xobs"] = np.asarray(st.x_obs_seq, dtype=np.float64)
    print(t, "done", flush=True)
# ---- SIR (sde/example_models/sir.py through the symnum stand-in): non-linear observation, inferred noise scale,
# one block of all observations
import torch  # noqa: E402
from tests.test_gpu_sir import make_sir_problem  # noqa: E402

prob = make_sir_problem(6, 4, 6, n_chains=2)
sir_r = R.load_models()[1]
sysr = ref.ConditionedDiffusionConstrainedSystem(
    prob["obs_interval"], prob["S"], prob["R"], torch.as_tensor(prob["y"]), 5, 3, 3, sir_r.forward_func,
    sir_r.generate_x_0, sir_r.generate_z, sir_r.obs_func, sir_r.generate_σ_y, False, dim_v_0=1)
rng = np.random.default_rng(23)
p_raw = rng.standard_normal(prob["q"].shape)
SIR_DT = 0.01
out.update(sir_T=prob["T"], sir_S=prob["S"], sir_obs_interval=prob["obs_interval"], sir_y=prob["y"], sir_q0=prob["q"],
           sir_xobs=prob["xobs"], sir_p_raw=p_raw, sir_dt=SIR_DT)
cs, lds, gs, ps = [], [], [], []
for i in range(prob["q"].shape[0]):
    st = ref.ConditionedDiffusionHamiltonianState(pos=prob["q"][i].copy(), x_obs_seq=prob["xobs"][i], partition=0)
    cs.append(np.asarray(sysr.constr(st), dtype=np.float64)); lds.append(float(sysr.log_det_sqrt_gram(st)))
    gs.append(np.asarray(sysr.grad_log_det_sqrt_gram(st), dtype=np.float64))
    ps.append(np.asarray(sysr.project_onto_cotangent_space(p_raw[i].copy(), st), dtype=np.float64))
out.update(sir_c=np.stack(cs), sir_ld=np.array(lds), sir_grad_ld=np.stack(gs), sir_p0=np.stack(ps))
for solver, wrapper in (("quasi_newton", ref.jitted_solve_projection_onto_manifold_quasi_newton),
                        ("newton", ref.jitted_solve_projection_onto_manifold_newton)):
    integ = ConstrainedLeapfrogIntegrator(sysr, step_size=SIR_DT, n_inner_step=1, reverse_check_tol=2e-8,
                                          projection_solver=wrapper, projection_solver_kwargs=TOL)
    qs, pps = [], []
    for i in range(prob["q"].shape[0]):
        st = ref.ConditionedDiffusionHamiltonianState(pos=prob["q"][i].copy(), x_obs_seq=prob["xobs"][i], partition=0,
                                                      mom=ps[i].copy())
        st = integ.step(st)
        qs.append(np.asarray(st.pos, dtype=np.float64)); pps.append(np.asarray(st.mom, dtype=np.float64))
    out.update({f"sir_{solver}_q": np.stack(qs), f"sir_{solver}_p": np.stack(pps)})
print("sir done", flush=True)
path = os.path.join(ROOT, "tests", "golden", "reference_pin_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
