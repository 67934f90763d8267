"""GPU parity: CUDA path (through the C ABI) vs the float64 autodiff oracle on small seeded cases.

Tolerances: the north-star asks for 1e-9 relative on positions / constraint residuals /
log-densities; most quantities agree to ~1e-12.
"""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.helpers import make_batched, make_fhn_problem

pytestmark = pytest.mark.gpu

# the last case has blocks of 10 observations (11 constraint rows): the 16-row FHN instantiation (mmd_ops_fhn_r16.cu),
# the R = 10 point of the reference's operation-time sweep (run_fhn_model_noiseless_obs_experiments.sh:16-22)
CASES = [(10, 5, 5), (12, 4, 5), (7, 6, 3), (20, 5, 10)]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: "T%d_S%d_R%d" % c)
def prob(request):
    T, S, R = request.param
    return make_fhn_problem(T, S, R, n_chains=3, nd=200)


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("part", [0, 1])
@pytest.mark.parametrize("off_manifold", [False, True])
def test_constr_logdet_grad(prob, part, off_manifold):
    rng = np.random.default_rng(1)
    q = prob["q"] + (0.05 * rng.standard_normal(prob["q"].shape) if off_manifold else 0.0)
    bc = make_batched(prob)
    bc.set_state(q, prob["xobs"], part)
    c = bc.constr()
    bc.linearize(True)
    ld = bc.log_det_sqrt_gram()
    g = bc.grad_log_det_sqrt_gram()
    sysm = prob["system"]
    for i in range(q.shape[0]):
        c_o = sysm._constr(torch.tensor(q[i]), torch.tensor(prob["xobs"][i]), part).numpy()
        assert c.shape[1] == c_o.shape[0]
        assert np.max(np.abs(c[i] - c_o)) < 1e-12
        pt = sysm.point(q[i], prob["xobs"][i], part)
        assert abs(ld[i] - pt["ld"]) < 1e-10 * max(1.0, abs(pt["ld"]))
        assert _rel(g[i], pt["grad_ld"].numpy()) < 1e-9
    bc.close()


@pytest.mark.parametrize("part", [0, 1])
def test_normal_space_component_and_xobs(prob, part):
    rng = np.random.default_rng(2)
    q = prob["q"]
    vct = rng.standard_normal(q.shape)
    bc = make_batched(prob)
    bc.set_state(q, prob["xobs"], part)
    bc.linearize(False)
    nsc = bc.normal_space_component(vct)
    sysm = prob["system"]
    for i in range(q.shape[0]):
        pt = sysm.point(q[i], prob["xobs"][i], part)
        nsc_o = sysm._normal_space_component(torch.tensor(vct[i]), pt["jac"], pt["chol"]).numpy()
        assert _rel(nsc[i], nsc_o) < 1e-9
    bc.update_x_obs_seq()
    _, _, x = bc.get_state()
    for i in range(q.shape[0]):
        x_o = sysm._generate_x_obs_seq(torch.tensor(q[i])).numpy()
        assert np.max(np.abs(x[i] - x_o)) < 1e-12
    bc.close()


@pytest.mark.parametrize("part", [0, 1])
def test_leapfrog_steps(prob, part):
    """Positions, momenta, Hamiltonian and iteration counts over 3 leapfrog steps given identical
    momenta (north-star correctness criterion)."""
    sysm = prob["system"]
    q0, xo = prob["q"], prob["xobs"]
    n = q0.shape[0]
    dt = 0.05
    rng = np.random.default_rng(3)
    p_raw = rng.standard_normal(q0.shape)
    bc = make_batched(prob)
    bc.set_state(q0, xo, part, p=p_raw)
    bc.linearize(True)
    bc.project_momentum()
    _, p_gpu0, _ = bc.get_state()
    h_gpu = [bc.hamiltonian()]
    traj = []
    for s in range(3):
        bc.leapfrog_step(dt)
        info = bc.step_info()
        qg, pg, _ = bc.get_state()
        traj.append((qg, pg, info, bc.hamiltonian()))
    for i in range(n):
        pt = sysm.point(q0[i], xo[i], part)
        p = sysm.project_onto_cotangent_space(torch.tensor(p_raw[i]), pt)
        assert _rel(p_gpu0[i], p.numpy()) < 1e-9
        assert abs(h_gpu[0][i] - sysm.h(torch.tensor(q0[i]), p, pt)) < 1e-9 * abs(h_gpu[0][i])
        q = torch.tensor(q0[i])
        for s in range(3):
            q, p, pt, inf = O.leapfrog_step(sysm, q, p, xo[i], part, dt, pt=pt)
            qg, pg, info, hg = traj[s]
            assert info["status"][i] == 0
            assert info["iters_fwd"][i] == inf["n_fwd"] and info["iters_rev"][i] == inf["n_back"]
            assert _rel(qg[i], q.numpy()) < 1e-9
            assert _rel(pg[i], p.numpy()) < 1e-8
            assert abs(hg[i] - sysm.h(q, p, pt)) < 1e-9 * abs(hg[i])
            assert abs(info["rev_dist"][i] - inf["rev_diff"]) < 1e-9
    bc.close()


def test_failed_step_keeps_state(prob):
    """A step that cannot converge leaves the chain where it was and reports why."""
    q0, xo = prob["q"], prob["xobs"]
    rng = np.random.default_rng(4)
    bc = make_batched(prob)
    bc.set_state(q0, xo, 0, p=rng.standard_normal(q0.shape))
    bc.linearize(True)
    bc.project_momentum()
    qa, pa, _ = bc.get_state()
    bc.opts.max_iters = 1
    bc.leapfrog_step(0.3)
    info = bc.step_info()
    assert np.all(info["status"] != 0)
    qb, pb, _ = bc.get_state()
    assert np.array_equal(qa, qb) and np.array_equal(pa, pb)
    bc.close()


def test_linear_interpolation_initialiser_values(prob):
    """k_init_interp against the oracle's find_initial_state_by_linear_interpolation (mici_extensions.py:1503-1526) on
    the same u, v_0 and x_obs_seq: the noise sequence itself, not only the constraint residual."""
    q_ref, xo = prob["q"], prob["xobs"]            # produced by the oracle's initialiser (tests/helpers.py)
    bc = make_batched(prob)
    bc.init_linear_interpolation(q_ref[:, :4], q_ref[:, 4:6], xo, 0)
    q, _, x = bc.get_state()
    assert np.array_equal(x, xo)
    assert np.max(np.abs(q - q_ref)) <= 1e-9 * max(1.0, np.max(np.abs(q_ref)))
    assert np.max(np.abs(bc.constr())) < 1e-8
    bc.close()
