"""GPU parity against vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/reference_pin_golden.npz, generated
in the build container by tests/golden/make_golden_reference_pin.py from /root/reference/sde/mici_extensions.py run
unmodified through a torch-backed jax stand-in): constraint, log-det, its gradient, projected momentum, Hamiltonian and
two constrained leapfrog steps with both projection solvers, both partitions; noiseless / Gaussian splitting / fixed
and inferred observation noise.  Tolerances as in the other parity tests (north-star: 1e-9 relative)."""
import os

import numpy as np
import pytest

from manifold_mcmc_for_diffusions_b200 import BatchedChains

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pin_golden.npz")


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("tag", ["noiseless", "ragged_gauss", "fixed_noise", "inferred_noise"])
@pytest.mark.parametrize("part", [0, 1])
def test_cuda_path_equals_the_reference_source(gold, tag, part):
    g = gold
    noise = int(g[f"{tag}_noise"])
    q0, xo, p_raw = g[f"{tag}_q0"], g[f"{tag}_xobs"], g[f"{tag}_p_raw"]
    n = q0.shape[0]
    mk = lambda: BatchedChains("fhn", 0.2, int(g[f"{tag}_S"]), int(g[f"{tag}_R"]), g[f"{tag}_y"], 5 if noise == 2 else 4, n,  # noqa: E731
                               noise=noise, sigma_fixed=float(g[f"{tag}_sigma"]),
                               use_gaussian_splitting=bool(g[f"{tag}_gaussian"]))
    bc = mk()
    bc.set_state(q0, xo, part, p=p_raw)
    c = bc.constr()
    bc.linearize(True)
    assert np.max(np.abs(c - g[f"{tag}_p{part}_c"])) < 1e-12
    ld = bc.log_det_sqrt_gram()
    assert np.max(np.abs(ld - g[f"{tag}_p{part}_ld"])) < 1e-10 * max(1.0, np.max(np.abs(ld)))
    for i in range(n):
        assert _rel(bc.grad_log_det_sqrt_gram()[i], g[f"{tag}_p{part}_grad_ld"][i]) < 1e-9
    bc.project_momentum()
    _, p0, _ = bc.get_state()
    assert _rel(p0, g[f"{tag}_p{part}_p0"]) < 1e-9
    h0 = bc.hamiltonian()
    assert np.max(np.abs(h0 - g[f"{tag}_p{part}_h0"])) < 1e-9 * np.max(np.abs(h0))
    bc.close()
    dt, n_steps = float(g["dt"]), int(g["n_steps"])
    for solver, name in ((0, "quasi_newton"), (1, "newton")):
        bc = mk()
        bc.opts.solver = solver
        bc.set_state(q0, xo, part, p=p_raw)
        bc.linearize(True)
        bc.project_momentum()
        for _ in range(n_steps):
            bc.leapfrog_step(dt)
            assert (bc.step_info()["status"] == 0).all()
        q, p, _ = bc.get_state()
        for i in range(n):
            assert _rel(q[i], g[f"{tag}_p{part}_{name}_q"][i]) < 1e-9
            assert _rel(p[i], g[f"{tag}_p{part}_{name}_p"][i]) < 1e-8
        h = bc.hamiltonian()
        assert np.max(np.abs(h - g[f"{tag}_p{part}_{name}_h"])) < 1e-9 * np.max(np.abs(h))
        bc.close()


@pytest.mark.parametrize("tag", ["noiseless", "inferred_noise"])
def test_partition_switch_equals_the_reference_source(gold, tag):
    g = gold
    noise = int(g[f"{tag}_noise"])
    bc = BatchedChains("fhn", 0.2, int(g[f"{tag}_S"]), int(g[f"{tag}_R"]), g[f"{tag}_y"], 5 if noise == 2 else 4, 1,
                       noise=noise, sigma_fixed=float(g[f"{tag}_sigma"]))
    bc.set_state(g[f"{tag}_q0"][:1], g[f"{tag}_xobs"][:1], 0)
    bc.switch_partition()
    _, _, x = bc.get_state()
    assert bc.partition == 1
    assert _rel(x[0], g[f"{tag}_switch_xobs"]) < 1e-12
    bc.close()


@pytest.mark.parametrize("solver", ["quasi_newton", "newton"])
@pytest.mark.parametrize("tag", ["noiseless", "inferred_noise"])
def test_drop_in_surface_call_counts_equal_the_reference(gold, tag, solver):
    """The drop-in module (mici_extensions.py of this package: same names as sde.mici_extensions) driven by the Mici
    step order books the same per-method call counts on the state as the reference's own source does for the same
    step -- the cost accounting the paper's cost-per-ESS figures rest on -- and lands on the same state."""
    from manifold_mcmc_for_diffusions_b200 import example_models, mici_extensions as me
    from manifold_mcmc_for_diffusions_b200.mici_compat.integrators import ConstrainedLeapfrogIntegrator

    g, m = gold, example_models.fhn
    noise = int(g[f"{tag}_noise"])
    system = me.ConditionedDiffusionConstrainedSystem(
        0.2, int(g[f"{tag}_S"]), int(g[f"{tag}_R"]), g[f"{tag}_y"], 5 if noise == 2 else 4, m.dim_x, m.dim_v, m.forward_func,
        m.generate_x_0, m.generate_z, m.obs_func, generate_σ=(None if noise == 0 else m.generate_σ_y),
        use_gaussian_splitting=bool(g[f"{tag}_gaussian"]), dim_v_0=m.dim_v_0)
    wrapper = (me.jitted_solve_projection_onto_manifold_quasi_newton if solver == "quasi_newton"
               else me.jitted_solve_projection_onto_manifold_newton)
    integ = ConstrainedLeapfrogIntegrator(
        system, step_size=float(g["dt"]), n_inner_step=1, reverse_check_tol=2e-8, projection_solver=wrapper,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10, max_iters=50))
    part = 0
    st = me.ConditionedDiffusionHamiltonianState(pos=g[f"{tag}_q0"][0].copy(), x_obs_seq=g[f"{tag}_xobs"][0], partition=part,
                                                 mom=g[f"{tag}_p{part}_p0"][0].copy())
    st = integ.step(st)
    counts = {k[2]: int(v) for k, v in st._call_counts.items()}
    ref_counts = dict(zip(g[f"{tag}_p{part}_{solver}_count_names"].tolist(), g[f"{tag}_p{part}_{solver}_count_values"].tolist()))
    assert counts == ref_counts, (counts, ref_counts)


def test_sir_cuda_path_equals_the_reference_source(gold):
    """The SIR model (sde/example_models/sir.py executed through the symnum stand-in inside the reference system):
    point quantities and one constrained leapfrog step with each solver."""
    g = gold
    q0, xo, p_raw = g["sir_q0"], g["sir_xobs"], g["sir_p_raw"]
    n, T = q0.shape[0], int(g["sir_T"])
    mk = lambda: BatchedChains("sir", float(g["sir_obs_interval"]), int(g["sir_S"]), T, g["sir_y"], 5, n, noise=2)  # noqa: E731
    bc = mk()
    bc.set_state(q0, xo, 0, p=p_raw)
    assert np.max(np.abs(bc.constr() - g["sir_c"])) < 1e-11 * max(1.0, float(np.max(np.abs(g["sir_y"]))))
    bc.linearize(True)
    assert np.max(np.abs(bc.log_det_sqrt_gram() - g["sir_ld"])) < 1e-10 * max(1.0, float(np.max(np.abs(g["sir_ld"]))))
    for i in range(n):
        assert _rel(bc.grad_log_det_sqrt_gram()[i], g["sir_grad_ld"][i]) < 1e-9
    bc.project_momentum()
    assert _rel(bc.get_state()[1], g["sir_p0"]) < 1e-9
    bc.close()
    for solver, name in ((0, "quasi_newton"), (1, "newton")):
        bc = mk()
        bc.opts.solver = solver
        bc.set_state(q0, xo, 0, p=p_raw)
        bc.linearize(True)
        bc.project_momentum()
        bc.leapfrog_step(float(g["sir_dt"]))
        assert (bc.step_info()["status"] == 0).all()
        q, p, _ = bc.get_state()
        for i in range(n):
            assert _rel(q[i], g[f"sir_{name}_q"][i]) < 1e-9 and _rel(p[i], g[f"sir_{name}_p"][i]) < 1e-8
        bc.close()
