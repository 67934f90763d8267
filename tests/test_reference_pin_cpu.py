"""Pins the restated oracle to the REFERENCE'S OWN SOURCE at trace level: `/root/reference/sde/mici_extensions.py` is
executed unmodified (oracle/reference_runner.py: torch-backed `jax` stand-in, minimal `mici` stand-in, the reference's
own model files through a SymPy-backed `symnum` stand-in) and compared, quantity by quantity, with `oracle/torch_oracle.py` on the same seeded inputs -- constraint,
log-det, its gradient, the cotangent projection, the Hamiltonian, and whole constrained leapfrog steps driven by the
Mici step order with the reference's projection-solver wrappers (positions, momenta, iteration counts).  CPU only; the
reference checkout exists only in the build container (skipped elsewhere; the GPU side compares with the vectors
generated from the same run, tests/golden/reference_pin_golden.npz)."""
import numpy as np
import pytest
import torch

from oracle import reference_runner as R
from oracle import torch_oracle as O
from tests.helpers import make_fhn_problem

pytestmark = pytest.mark.skipif(not R.available(), reason="reference checkout not present")

CASES = [dict(T=10, S=5, R=5, noise=0, gaussian=False), dict(T=12, S=4, R=5, noise=0, gaussian=True),
         dict(T=10, S=5, R=5, noise=1, gaussian=False), dict(T=10, S=5, R=5, noise=2, gaussian=False)]


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module", params=CASES, ids=lambda c: "T%d_S%d_R%d_noise%d_%s" % (
    c["T"], c["S"], c["R"], c["noise"], "gauss" if c["gaussian"] else "std"))
def setup(request):
    c = request.param
    prob = make_fhn_problem(c["T"], c["S"], c["R"], n_chains=1, nd=200, noise=c["noise"], gaussian=c["gaussian"])
    ref = R.load()
    sysr = R.make_fhn_system(0.2, c["S"], c["R"], prob["y"], noise=c["noise"], sigma=prob["sigma"],
                             use_gaussian_splitting=c["gaussian"])
    return prob, ref, sysr


@pytest.mark.parametrize("part", [0, 1])
def test_point_quantities_equal_the_reference(setup, part):
    prob, ref, sysr = setup
    syso = prob["system"]
    rng = np.random.default_rng(7)
    for i in range(prob["q"].shape[0]):
        for off in (0.0, 0.03):
            q = prob["q"][i] + off * rng.standard_normal(prob["q"][i].shape)
            xo = prob["xobs"][i]
            st = ref.ConditionedDiffusionHamiltonianState(pos=q, x_obs_seq=xo, partition=part,
                                                          mom=rng.standard_normal(q.shape))
            pt = syso.point(q, xo, part)
            c_r, c_o = np.asarray(sysr.constr(st)), syso._constr(torch.tensor(q), torch.tensor(xo), part).numpy()
            assert np.max(np.abs(c_r - c_o)) < 1e-13      # (the two step maps differ by rounding: 1e-16 per step)
            assert abs(float(sysr.log_det_sqrt_gram(st)) - float(pt["ld"])) < 1e-11 * max(1.0, abs(float(pt["ld"])))
            assert _rel(sysr.grad_log_det_sqrt_gram(st), pt["grad_ld"]) < 1e-10
            vct = rng.standard_normal(q.shape)
            p_ref = sysr.project_onto_cotangent_space(vct.copy(), st)
            p_orc = syso.project_onto_cotangent_space(torch.tensor(vct), pt)
            assert _rel(p_ref, p_orc) < 1e-11
            st.mom = np.asarray(p_ref)
            assert abs(float(sysr.h(st)) - float(syso.h(torch.tensor(q), p_orc, pt))) < 1e-11 * abs(float(sysr.h(st)))


@pytest.mark.parametrize("solver", ["quasi_newton", "newton"])
@pytest.mark.parametrize("part", [0, 1])
def test_leapfrog_steps_equal_the_reference(setup, part, solver):
    """Mici's ConstrainedLeapfrogIntegrator step order over the REFERENCE system and the REFERENCE's projection-solver
    wrapper against the oracle's leapfrog_step: same positions, momenta and solver iteration counts."""
    from manifold_mcmc_for_diffusions_b200.mici_compat.integrators import ConstrainedLeapfrogIntegrator

    prob, ref, sysr = setup
    syso = prob["system"]
    wrapper = (ref.jitted_solve_projection_onto_manifold_quasi_newton if solver == "quasi_newton"
               else ref.jitted_solve_projection_onto_manifold_newton)
    tol = dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10, max_iters=50)
    integ = ConstrainedLeapfrogIntegrator(sysr, step_size=0.05, n_inner_step=1, reverse_check_tol=2e-8,
                                          projection_solver=wrapper, projection_solver_kwargs=tol)
    rng = np.random.default_rng(9)
    for i in range(prob["q"].shape[0]):
        q0, xo = prob["q"][i], prob["xobs"][i]
        st = ref.ConditionedDiffusionHamiltonianState(pos=q0.copy(), x_obs_seq=xo, partition=part)
        p_raw = rng.standard_normal(q0.shape)
        st.mom = np.asarray(sysr.project_onto_cotangent_space(p_raw.copy(), st))
        pt = syso.point(q0, xo, part)
        q, p = torch.tensor(q0), syso.project_onto_cotangent_space(torch.tensor(p_raw), pt)
        for s in range(2):
            st = integ.step(st)
            q, p, pt, inf = O.leapfrog_step(syso, q, p, xo, part, 0.05, pt=pt, solver=solver, **tol)
            assert _rel(st.pos, q) < 1e-11
            assert _rel(st.mom, p) < 1e-9
            assert abs(float(sysr.h(st)) - float(syso.h(q, p, pt))) < 1e-10 * abs(float(syso.h(q, p, pt)))


def test_partition_switch_and_initialiser_equal_the_reference(setup):
    prob, ref, sysr = setup
    syso = prob["system"]
    q, xo = prob["q"][0], prob["xobs"][0]
    st = ref.ConditionedDiffusionHamiltonianState(pos=q, x_obs_seq=xo, partition=0)
    ref.SwitchPartitionTransition(sysr).sample(st, None)
    assert st.partition == 1
    x_o = syso.generate_x_obs_seq(torch.tensor(q)) if hasattr(syso, "generate_x_obs_seq") else None
    if x_o is not None:
        assert _rel(st.x_obs_seq, x_o) < 1e-13


def test_sir_system_equals_the_reference():
    """The SIR model (Euler-Maruyama on the log-transformed SDE, non-linear observation, inferred noise scale, one block
    of all observations) through the reference's own source against the oracle: point quantities and one leapfrog step
    with each solver."""
    import torch as _torch

    from manifold_mcmc_for_diffusions_b200.mici_compat.integrators import ConstrainedLeapfrogIntegrator
    from oracle.models import sir
    from tests.test_gpu_sir import make_sir_problem

    prob = make_sir_problem(6, 4, 6, n_chains=1)
    ref, syso = R.load(), prob["system"]
    sir_r = R.load_models()[1]          # the reference's own sde/example_models/sir.py
    sysr = ref.ConditionedDiffusionConstrainedSystem(
        prob["obs_interval"], prob["S"], prob["R"], _torch.as_tensor(prob["y"]), 5, 3, 3, sir_r.forward_func,
        sir_r.generate_x_0, sir_r.generate_z, sir_r.obs_func, sir_r.generate_σ_y, False, dim_v_0=1)
    q0, xo = prob["q"][0], prob["xobs"][0]
    rng = np.random.default_rng(3)
    st = ref.ConditionedDiffusionHamiltonianState(pos=q0.copy(), x_obs_seq=xo, partition=0)
    pt = syso.point(q0, xo, 0)
    # (observations are counts of a few hundred: 1e-11 absolute is 1e-14 relative)
    assert np.max(np.abs(np.asarray(sysr.constr(st)) - syso._constr(_torch.tensor(q0), _torch.tensor(xo), 0).numpy())) < 1e-11
    assert abs(float(sysr.log_det_sqrt_gram(st)) - float(pt["ld"])) < 1e-11 * max(1.0, abs(float(pt["ld"])))
    assert _rel(sysr.grad_log_det_sqrt_gram(st), pt["grad_ld"]) < 1e-10
    p_raw = rng.standard_normal(q0.shape)
    tol = dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10, max_iters=50)
    for solver, wrapper in (("quasi_newton", ref.jitted_solve_projection_onto_manifold_quasi_newton),
                            ("newton", ref.jitted_solve_projection_onto_manifold_newton)):
        integ = ConstrainedLeapfrogIntegrator(sysr, step_size=0.01, n_inner_step=1, reverse_check_tol=2e-8,
                                              projection_solver=wrapper, projection_solver_kwargs=tol)
        st = ref.ConditionedDiffusionHamiltonianState(pos=q0.copy(), x_obs_seq=xo, partition=0)
        st.mom = np.asarray(sysr.project_onto_cotangent_space(p_raw.copy(), st))
        p = syso.project_onto_cotangent_space(_torch.tensor(p_raw), pt)
        assert _rel(st.mom, p) < 1e-11
        st = integ.step(st)
        q, p, _, _ = O.leapfrog_step(syso, q0, p, xo, 0, 0.01, pt=pt, solver=solver, **tol)
        assert _rel(st.pos, q) < 1e-11 and _rel(st.mom, p) < 1e-9


@pytest.mark.parametrize("gaussian", [False, True])
def test_hmc_target_and_initialiser_equal_the_reference(gaussian):
    """The standard-HMC baseline target (conditioned_diffusion_neg_log_dens_and_grad, :82-205) and the
    linear-interpolation initialiser (:1479-1547) of the reference source against the oracle."""
    import torch as _torch

    from oracle.models import fhn

    prob = make_fhn_problem(8, 5, 4, n_chains=1, nd=100, noise=2, gaussian=gaussian)
    ref, syso = R.load(), prob["system"]
    y = _torch.as_tensor(prob["y"])
    args = (0.2, 5, y, 5, 2, 2, fhn.forward_func, fhn.generate_x_0, fhn.generate_z, fhn.generate_σ_y, fhn.obs_func, gaussian)
    fhn_r = R.load_models()[0]
    nld_r, grad_r = ref.conditioned_diffusion_neg_log_dens_and_grad(
        0.2, 5, y, 5, 2, 2, fhn_r.forward_func, fhn_r.generate_x_0, fhn_r.generate_z, fhn_r.generate_σ_y, fhn_r.obs_func,
        gaussian)
    nld_o, vg_o = O.conditioned_diffusion_neg_log_dens_and_grad(*args)
    rng = np.random.default_rng(2)
    q = np.concatenate([0.3 * rng.standard_normal(4), [np.log(0.1)], prob["q"][0][5:7], prob["q"][0][7:7 + 8 * 5 * 2]])
    g_r, v_r = grad_r(q)
    v_o, g_o = vg_o(q)
    assert abs(v_r - float(v_o)) < 1e-12 * abs(v_r) and abs(nld_r(q) - float(nld_o(q))) < 1e-12 * abs(v_r)
    assert _rel(g_r, g_o) < 1e-11
    # initialiser: same u, v_0 and x_obs_seq on both sides
    sysr = R.make_fhn_system(0.2, 5, 4, prob["y"], noise=2, use_gaussian_splitting=gaussian)
    x_init = np.concatenate((prob["y"], 0.5 * np.random.default_rng(5).standard_normal(prob["y"].shape)), -1)
    u, v0 = 0.4 * rng.standard_normal(5), rng.standard_normal(2)
    st = ref.find_initial_state_by_linear_interpolation(sysr, np.random.default_rng(0), lambda r: x_init,
                                                        u=_torch.tensor(u), v_0=_torch.tensor(v0))
    q_o, x_o = O.find_initial_state_by_linear_interpolation(syso, np.random.default_rng(0), lambda r: x_init, u=u, v_0=v0)
    assert _rel(st.pos, q_o) < 1e-11 and _rel(st.x_obs_seq, x_o) < 1e-14
    assert np.max(np.abs(np.asarray(sysr.constr(st)))) < 1e-9


def test_model_definitions_equal_the_reference():
    """The reference's own model files (sde/example_models/fhn.py, sir.py with sde/integrators.py and
    sde/transforms.py, executed through the SymPy-backed symnum stand-in) against oracle/models.py: step maps, their
    Jacobians, generators and observation functions; SIR also inside the -500 clip region (sir.py:54-70)."""
    import torch as _torch

    from oracle.models import fhn as fhn_o, sir as sir_o

    fhn_r, sir_r = R.load_models()
    rng = np.random.default_rng(0)
    for _ in range(5):
        u = _torch.tensor(rng.standard_normal(5))
        assert _rel(fhn_r.generate_z(u), fhn_o.generate_z(u)) < 1e-15 and _rel(sir_r.generate_z(u), sir_o.generate_z(u)) < 1e-15
        assert _rel(fhn_r.generate_σ_y(u), fhn_o.generate_σ_y(u)) < 1e-15
        z = fhn_o.generate_z(0.5 * u)
        x, v, v0 = (_torch.tensor(rng.standard_normal(2)) for _ in range(3))
        assert _rel(fhn_r.generate_x_0(z, v0), fhn_o.generate_x_0(z, v0)) < 1e-15
        f_r = lambda z_, x_, v_: fhn_r.forward_func(z_, x_, v_, 0.008)  # noqa: E731
        f_o = lambda z_, x_, v_: fhn_o.forward_func(z_, x_, v_, 0.008)  # noqa: E731
        assert _rel(f_r(z, x, v), f_o(z, x, v)) < 1e-14
        for a, b in zip(_torch.func.jacrev(f_r, argnums=(0, 1, 2))(z, x, v), _torch.func.jacrev(f_o, argnums=(0, 1, 2))(z, x, v)):
            assert _rel(a, b) < 1e-12
        zs = sir_o.generate_z(0.3 * u)
        for xs in (_torch.tensor([np.log(700.0), np.log(5.0), 0.3]) + 0.1 * _torch.tensor(rng.standard_normal(3)),
                   _torch.tensor([np.log(700.0), -600.0, 0.3]), _torch.tensor([-501.0, np.log(3.0), -0.2])):
            vs = _torch.tensor(rng.standard_normal(3))
            g_r = lambda z_, x_, v_: sir_r.forward_func(z_, x_, v_, 0.05)  # noqa: E731
            g_o = lambda z_, x_, v_: sir_o.forward_func(z_, x_, v_, 0.05)  # noqa: E731
            assert _rel(g_r(zs, xs, vs), g_o(zs, xs, vs)) < 1e-14
            for a, b in zip(_torch.func.jacrev(g_r, argnums=(0, 1, 2))(zs, xs, vs),
                            _torch.func.jacrev(g_o, argnums=(0, 1, 2))(zs, xs, vs)):
                assert _rel(a, b) < 1e-12 or float(_torch.max(_torch.abs(a - b))) < 1e-12
        xx = _torch.tensor(rng.standard_normal((4, 3)))
        assert _rel(sir_r.obs_func(xx), sir_o.obs_func(xx)) < 1e-15 and _rel(fhn_r.obs_func(xx[:, :2]), fhn_o.obs_func(xx[:, :2])) < 1e-15
        assert _rel(sir_r.generate_x_0(zs, v0[:1]), sir_o.generate_x_0(zs, v0[:1])) < 1e-15


def test_canonical_size_golden_equals_the_reference():
    """The committed canonical-size vectors (tests/golden/fhn_T100_S25_R5_golden.npz: FHN noiseless T=100, S=25, R=5, the
    shape of the bench workload; frozen from the oracle) against the reference's own source at that size: log-det,
    its gradient, and two constrained leapfrog steps (chain 0 / partition 0 with the reference's quasi-Newton wrapper)."""
    import os

    from manifold_mcmc_for_diffusions_b200.mici_compat.integrators import ConstrainedLeapfrogIntegrator

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fhn_T100_S25_R5_golden.npz"),
                allow_pickle=True)
    ref = R.load()
    sysr = R.make_fhn_system(0.2, int(g["S"]), int(g["R"]), g["y"])
    c = 0
    st = ref.ConditionedDiffusionHamiltonianState(pos=g["q0"][c].copy(), x_obs_seq=g["xobs"][c], partition=c % 2)
    assert abs(float(sysr.log_det_sqrt_gram(st)) - float(g["ld"][c])) < 1e-10 * abs(float(g["ld"][c]))
    assert _rel(sysr.grad_log_det_sqrt_gram(st), g["grad_ld"][c]) < 1e-9
    st.mom = np.asarray(sysr.project_onto_cotangent_space(g["p_raw"][c].copy(), st))
    integ = ConstrainedLeapfrogIntegrator(
        sysr, step_size=float(g["dt"]), n_inner_step=1, reverse_check_tol=2e-8,
        projection_solver=ref.jitted_solve_projection_onto_manifold_quasi_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10, max_iters=50))
    for s in range(2):
        st = integ.step(st)
        assert _rel(st.pos, g["traj_q"][c][s]) < 1e-10 and _rel(st.mom, g["traj_p"][c][s]) < 1e-8
        assert abs(float(sysr.h(st)) - float(g["traj_h"][c][s + 1])) < 1e-10 * abs(float(g["traj_h"][c][s + 1]))


@pytest.mark.parametrize("tag", ["fhn_noisy", "sir"])
def test_bundled_data_goldens_equal_the_reference(tag):
    """The committed vectors on the reference's bundled data sets (tests/golden/bundled_configs_golden.npz, BASELINE
    configs 2 and 3: FHN noisy observations T=100 S=40 R=5, SIR boarding-school data) against the reference's own source:
    constraint, log-det, its gradient and one constrained leapfrog step with the Newton solver (the scripts' default)."""
    import os

    import torch as _torch

    from manifold_mcmc_for_diffusions_b200.mici_compat.integrators import ConstrainedLeapfrogIntegrator

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bundled_configs_golden.npz"))
    ref = R.load()
    fhn_r, sir_r = R.load_models()
    y = _torch.as_tensor(np.asarray(g[f"{tag}_y"], dtype=np.float64))
    if tag == "fhn_noisy":
        sysr = ref.ConditionedDiffusionConstrainedSystem(
            float(g[f"{tag}_obs_interval"]), int(g[f"{tag}_S"]), int(g[f"{tag}_R"]), y, 4, 2, 2, fhn_r.forward_func,
            fhn_r.generate_x_0, fhn_r.generate_z, fhn_r.obs_func, float(g[f"{tag}_sigma"]), False, dim_v_0=2)
    else:
        sysr = ref.ConditionedDiffusionConstrainedSystem(
            float(g[f"{tag}_obs_interval"]), int(g[f"{tag}_S"]), int(g[f"{tag}_T"]), y, 4, 3, 3, sir_r.forward_func,
            sir_r.generate_x_0, sir_r.generate_z, sir_r.obs_func, float(g[f"{tag}_sigma"]), False, dim_v_0=1)
    st = ref.ConditionedDiffusionHamiltonianState(pos=g[f"{tag}_q0"].copy(), x_obs_seq=g[f"{tag}_xobs"], partition=0)
    assert np.max(np.abs(np.asarray(sysr.constr(st)) - g[f"{tag}_c"])) < 1e-11 * max(1.0, float(np.max(np.abs(g[f"{tag}_y"]))))
    ld = float(g[f"{tag}_ld"])
    assert abs(float(sysr.log_det_sqrt_gram(st)) - ld) < 1e-10 * max(1.0, abs(ld))
    assert _rel(sysr.grad_log_det_sqrt_gram(st), g[f"{tag}_grad_ld"]) < 1e-9
    st.mom = np.asarray(sysr.project_onto_cotangent_space(g[f"{tag}_p_raw"].copy(), st))
    integ = ConstrainedLeapfrogIntegrator(
        sysr, step_size=float(g[f"{tag}_dt"]), n_inner_step=1, reverse_check_tol=2e-8,
        projection_solver=ref.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10, max_iters=50))
    st = integ.step(st)
    assert _rel(st.pos, g[f"{tag}_newton_q"]) < 1e-10 and _rel(st.mom, g[f"{tag}_newton_p"]) < 1e-8
    h_ref = float(g[f"{tag}_newton_h"])
    assert abs(float(sysr.h(st)) - h_ref) < 1e-9 * max(1.0, abs(h_ref))


def test_metric_adapter_equals_the_reference():
    """OnlineBlockDiagonalMetricAdapter (:1804-1931): this package's class against the reference's own, fed the same
    chains (single chain and the multi-chain combination in finalize)."""
    from manifold_mcmc_for_diffusions_b200 import mici_extensions as me

    class _State:
        def __init__(self, pos):
            self.pos = pos

    class _Tr:
        def __init__(self):
            self.system = type("S", (), {"metric": None})()

    def run(ad, draws):
        st = ad.initialize(_State(draws[0]), None)
        for d in draws:
            ad.update(st, _State(d), None, None)
        return st

    ref = R.load()
    rng = np.random.default_rng(4)
    chains = [rng.standard_normal((k, 8)) @ rng.standard_normal((8, 8)) + i for i, k in enumerate((25, 40, 13))]
    for states in (lambda ad: run(ad, chains[0]), lambda ad: [run(ad, c) for c in chains]):
        a_r, a_o = ref.OnlineBlockDiagonalMetricAdapter(5), me.OnlineBlockDiagonalMetricAdapter(5)
        t_r, t_o = _Tr(), _Tr()
        a_r.finalize(states(a_r), t_r)
        a_o.finalize(states(a_o), t_o)
        b_r, b_o = t_r.system.metric.blocks, t_o.system.metric.blocks
        assert _rel(b_o[0].array, b_r[0].array) < 1e-12
        assert type(b_o[1]).__name__ == type(b_r[1]).__name__ == "IdentityMatrix"
