"""The reference-facing plugin surface (manifold_mcmc_for_diffusions_b200.mici_extensions) driven the way
Mici drives sde.mici_extensions: system methods on a ChainState, ConstrainedLeapfrogIntegrator with the
device projection solver, NUTS transition + partition switch.  Checked against the float64 oracle."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.helpers import OBS_INTERVAL, make_fhn_problem

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _system(prob, **kw):
    from manifold_mcmc_for_diffusions_b200 import example_models, mici_extensions as me

    m = example_models.fhn
    gen_sigma = None if prob["noise"] == 0 else (prob["sigma"] if prob["noise"] == 1 else m.generate_σ_y)
    return me.ConditionedDiffusionConstrainedSystem(
        OBS_INTERVAL, prob["S"], prob["R"], prob["y"], prob["dim_u"], m.dim_x, m.dim_v, m.forward_func,
        m.generate_x_0, m.generate_z, m.obs_func, generate_σ=gen_sigma,
        use_gaussian_splitting=prob["gaussian"], dim_v_0=m.dim_v_0, **kw)


@pytest.fixture(scope="module", params=[dict(), dict(noise=2, sigma=0.2), dict(gaussian=True)],
                ids=["noiseless", "noisy_param", "gaussian"])
def prob(request):
    return make_fhn_problem(10, 5, 5, n_chains=2, nd=200, **request.param)


def _flat_blocks(blocks):
    out = []
    for b in blocks:
        if b is None:
            continue
        b = b.numpy() if hasattr(b, "numpy") else np.asarray(b)
        out += [b] if b.ndim == 2 else list(b)
    return out


@pytest.mark.parametrize("part", [0, 1])
def test_system_methods_match_oracle(prob, part):
    from manifold_mcmc_for_diffusions_b200 import mici_extensions as me

    system = _system(prob)
    sysm = prob["system"]
    assert system.num_partition == sysm.num_partition and system.dim_q == sysm.dim_q
    rng = np.random.default_rng(5)
    q = prob["q"][0] + 0.02 * rng.standard_normal(prob["q"][0].shape)
    state = me.ConditionedDiffusionHamiltonianState(pos=q, x_obs_seq=prob["xobs"][0], partition=part)
    pt = sysm.point(q, prob["xobs"][0], part)
    c_o = sysm._constr(torch.tensor(q), torch.tensor(prob["xobs"][0]), part).numpy()
    assert np.max(np.abs(system.constr(state) - c_o)) < 1e-12
    assert abs(system.log_det_sqrt_gram(state) - pt["ld"]) < 1e-10 * max(1, abs(pt["ld"]))
    assert _rel(system.grad_log_det_sqrt_gram(state), pt["grad_ld"].numpy()) < 1e-9
    mom = rng.standard_normal(q.shape)
    state.mom = mom.copy()
    nsc_o = sysm._normal_space_component(torch.tensor(mom), pt["jac"], pt["chol"]).numpy()
    assert _rel(system.normal_space_component(state, mom), nsc_o) < 1e-9
    assert abs(system.h(state) - sysm.h(torch.tensor(q), torch.tensor(mom), pt)) < 1e-9 * abs(system.h(state))
    # dense blocks rebuilt from the compressed factors vs the oracle's jacrev blocks / Cholesky factors
    dc_du, dc_dv, dc_dn = system.jacob_constr_blocks(state)
    for mine, ref in zip(dc_du, _flat_blocks(pt["jac"][0])):
        assert _rel(mine, ref) < 1e-9
    for mine, ref in zip(dc_dv, _flat_blocks(pt["jac"][1])):
        assert mine.shape == ref.shape and _rel(mine, ref) < 1e-9
    chol_C, chol_D = system.chol_gram_blocks(state)
    assert _rel(chol_C, pt["chol"][0].numpy()) < 1e-9
    for mine, ref in zip(chol_D, _flat_blocks(pt["chol"][1])):
        assert _rel(mine, ref) < 1e-8
    # call counts like Mici: one evaluation per cached method
    counts = {k[2]: v for k, v in state._call_counts.items()}
    assert counts["constr"] == 1 and counts["grad_log_det_sqrt_gram"] == 1


@pytest.mark.parametrize("solver", ["quasi_newton", "newton"])
@pytest.mark.parametrize("part", [0, 1])
def test_mici_integrator_steps_match_oracle(prob, part, solver):
    from manifold_mcmc_for_diffusions_b200 import mici_compat, mici_extensions as me

    system = _system(prob)
    sysm = prob["system"]
    integrator = mici_compat.integrators.ConstrainedLeapfrogIntegrator(
        system, step_size=0.05, n_inner_step=1, reverse_check_tol=2e-8,
        projection_solver=(me.jitted_solve_projection_onto_manifold_newton if solver == "newton"
                           else me.jitted_solve_projection_onto_manifold_quasi_newton),
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    rng = np.random.default_rng(3)
    q0, xo = prob["q"][1], prob["xobs"][1]
    p_raw = rng.standard_normal(q0.shape)
    state = me.ConditionedDiffusionHamiltonianState(pos=q0.copy(), x_obs_seq=xo, partition=part)
    state.mom = system.project_onto_cotangent_space(p_raw.copy(), state)
    pt = sysm.point(q0, xo, part)
    p = sysm.project_onto_cotangent_space(torch.tensor(p_raw), pt)
    q = torch.tensor(q0)
    for s in range(3):
        state = integrator.step(state)
        q, p, pt, inf = O.leapfrog_step(sysm, q, p, xo, part, 0.05, pt=pt, solver=solver)
        assert _rel(state.pos, q.numpy()) < 1e-9
        assert _rel(state.mom, p.numpy()) < 1e-8
        assert abs(system.h(state) - sysm.h(q, p, pt)) < 1e-9 * abs(system.h(state))
    key = [k for k in state._call_counts if k[2] == "constr"][0]
    assert state._call_counts[key] > 0   # projection iterations are booked on `constr` (:1382-1387)


def test_failed_projection_raises_convergence_error(prob):
    from manifold_mcmc_for_diffusions_b200 import mici_compat, mici_extensions as me

    system = _system(prob)
    state_prev = me.ConditionedDiffusionHamiltonianState(pos=prob["q"][0].copy(), x_obs_seq=prob["xobs"][0])
    state = state_prev.copy()
    state.mom = np.zeros_like(state.pos)
    state.pos = state.pos + 0.3 * np.random.default_rng(0).standard_normal(state.pos.shape)
    before = state.pos.copy()
    with pytest.raises(mici_compat.errors.ConvergenceError):
        me.jitted_solve_projection_onto_manifold_quasi_newton(state, state_prev, 0.3, system, max_iters=1)
    assert np.array_equal(state.pos, before)


def test_nuts_transitions_and_partition_switch():
    from manifold_mcmc_for_diffusions_b200 import mici_compat, mici_extensions as me

    prob = make_fhn_problem(10, 5, 5, n_chains=1, nd=200)
    system = _system(prob)
    integrator = mici_compat.integrators.ConstrainedLeapfrogIntegrator(
        system, step_size=0.05, projection_solver=me.jitted_solve_projection_onto_manifold_quasi_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    rng = np.random.default_rng(11)
    y = prob["y"]
    state = me.find_initial_state_by_linear_interpolation(
        system, rng, lambda r: np.concatenate((y, 0.5 * r.standard_normal(y.shape)), -1))
    assert np.max(np.abs(system.constr(state))) < 1e-8
    assert np.max(np.abs(system.normal_space_component(state, state.mom))) < 1e-9
    transitions = {
        "momentum": mici_compat.transitions.IndependentMomentumTransition(system),
        "integration": mici_compat.transitions.MultinomialDynamicIntegrationTransition(system, integrator, max_tree_depth=3),
        "switch_partition": me.SwitchPartitionTransition(system),
    }
    n_step = 0
    for it in range(4):
        for key, tr in transitions.items():
            state, st = tr.sample(state, rng)
            if key == "integration":
                n_step += st["n_step"]
                assert 0.0 <= st["accept_stat"] <= 1.0 and np.isfinite(st["hamiltonian"])
        assert state.partition == (it + 1) % 2
        assert np.max(np.abs(system.constr(state))) < 1e-7    # on the manifold of the NEW partition
    assert n_step >= 4


def test_sir_system_through_the_plugin_surface():
    """SIR model selected by the tagged callables of example_models.sir (non-linear observation, inferred
    noise scale through a callable generate_σ), one Mici-style leapfrog step with the Newton solver."""
    from manifold_mcmc_for_diffusions_b200 import example_models, mici_compat, mici_extensions as me
    from tests.test_gpu_sir import make_sir_problem

    prob = make_sir_problem(6, 4, 3, n_chains=1)
    m = example_models.sir
    system = me.ConditionedDiffusionConstrainedSystem(
        prob["obs_interval"], prob["S"], prob["R"], prob["y"], 5, m.dim_x, m.dim_v, m.forward_func, m.generate_x_0,
        m.generate_z, m.obs_func, generate_σ=m.generate_σ_y, dim_v_0=m.dim_v_0)
    sysm = prob["system"]
    q0, xo = prob["q"][0], prob["xobs"][0]
    state = me.ConditionedDiffusionHamiltonianState(pos=q0.copy(), x_obs_seq=xo, partition=1)
    pt = sysm.point(q0, xo, 1)
    assert np.max(np.abs(system.constr(state))) < 1e-9
    assert abs(system.log_det_sqrt_gram(state) - pt["ld"]) < 1e-10 * max(1, abs(pt["ld"]))
    assert _rel(system.grad_log_det_sqrt_gram(state), pt["grad_ld"].numpy()) < 1e-9
    rng = np.random.default_rng(2)
    p_raw = rng.standard_normal(q0.shape)
    state.mom = system.project_onto_cotangent_space(p_raw.copy(), state)
    integrator = mici_compat.integrators.ConstrainedLeapfrogIntegrator(
        system, step_size=0.02, projection_solver=me.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    new = integrator.step(state)
    p = sysm.project_onto_cotangent_space(torch.tensor(p_raw), pt)
    q, p, ptn, _ = O.leapfrog_step(sysm, q0, p, xo, 1, 0.02, pt=pt, solver="newton")
    assert _rel(new.pos, q.numpy()) < 1e-9 and _rel(new.mom, p.numpy()) < 1e-8
