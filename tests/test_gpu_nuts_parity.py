"""Device path of the batched NUTS transition against the recursive Mici-style transition
(mici_compat.transitions.MultinomialDynamicIntegrationTransition) over the CUDA-backed single-chain system
(mici_extensions.ConditionedDiffusionConstrainedSystem + ConstrainedLeapfrogIntegrator with the device projection
solver), chain by chain with the SAME momenta and the SAME uniforms: n_step, tree_depth, termination flags,
accept_stat and the selected proposal must agree."""

import numpy as np
import pytest

from tests.helpers import OBS_INTERVAL, make_batched, make_fhn_problem
from tests.test_nuts_host_logic import _RecordingRng, _ReplayRng

pytestmark = pytest.mark.gpu


def test_batched_nuts_device_path_equals_recursive_transition():
    from manifold_mcmc_for_diffusions_b200 import example_models, mici_compat, mici_extensions as me
    from manifold_mcmc_for_diffusions_b200.nuts import BatchedNUTS

    n, depth, eps = 12, 4, 0.06
    prob = make_fhn_problem(10, 5, 5, n_chains=n, nd=200)
    bc = make_batched(prob)
    bc.set_state(prob["q"], prob["xobs"], 0)
    for it in range(20):                                   # towards the typical set
        bc.hmc_transition(0.02, 4, 3, it)
    q_start, _, x_start = bc.get_state()
    part = bc.partition
    nuts = BatchedNUTS(bc, max_tree_depth=depth)
    rec, trace = _RecordingRng(9), {}
    stats = nuts.transition(eps, rec, 3, 100, switch_partition=False, trace=trace)
    q_end, _, _ = bc.get_state()
    q0, p0, x0 = trace["state0"]
    assert np.array_equal(q0, q_start) and trace["partition"] == part
    draws = np.array(rec.draws)

    m = example_models.fhn
    system = me.ConditionedDiffusionConstrainedSystem(
        OBS_INTERVAL, prob["S"], prob["R"], prob["y"], 4, m.dim_x, m.dim_v, m.forward_func, m.generate_x_0,
        m.generate_z, m.obs_func, dim_v_0=m.dim_v_0)
    integrator = mici_compat.integrators.ConstrainedLeapfrogIntegrator(
        system, step_size=eps, projection_solver=me.jitted_solve_projection_onto_manifold_quasi_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50), reverse_check_tol=2e-8)
    tr = mici_compat.transitions.MultinomialDynamicIntegrationTransition(system, integrator, max_tree_depth=depth)
    depths = set()
    for c in range(n):
        state = me.ConditionedDiffusionHamiltonianState(pos=q0[c].copy(), x_obs_seq=x0[c].copy(), partition=part,
                                                        mom=p0[c].copy())
        new, ref = tr.sample(state, _ReplayRng(draws[:, c]))
        assert ref["n_step"] == stats["n_step"][c], (c, ref, {k: v[c] for k, v in stats.items()})
        assert ref["tree_depth"] == stats["tree_depth"][c], c
        assert ref["convergence_error"] == bool(stats["convergence_error"][c]), c
        assert ref["non_reversible_step"] == bool(stats["non_reversible_step"][c]), c
        assert ref["diverging"] == bool(stats["diverging"][c]), c
        assert abs(ref["accept_stat"] - stats["accept_stat"][c]) < 1e-9, c
        assert np.max(np.abs(new.pos - q_end[c])) <= 1e-9 * np.max(np.abs(q_end[c])), c
        depths.add(int(ref["tree_depth"]))
    assert max(depths) >= 2 and (stats["n_step"] >= 3).any()
    bc.close()
