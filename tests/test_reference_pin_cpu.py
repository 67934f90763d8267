"""Pins the restated oracle to the REFERENCE'S OWN SOURCE at trace level: `/root/reference/sde/mici_extensions.py` is
executed unmodified (oracle/reference_runner.py: torch-backed `jax` stand-in, minimal `mici` stand-in, torch model
callables) and compared, quantity by quantity, with `oracle/torch_oracle.py` on the same seeded inputs -- constraint,
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
    prob = make_fhn_problem(c["T"], c["S"], c["R"], n_chains=2, nd=200, noise=c["noise"], gaussian=c["gaussian"])
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
            assert _rel(sysr.constr(st), syso._constr(torch.tensor(q), torch.tensor(xo), part)) < 1e-13 or \
                np.max(np.abs(np.asarray(sysr.constr(st)))) < 1e-13
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
