"""GPU parity of the Newton projection solver (newton_projection mici_extensions.py:1065-1135 with
lu_jacob_product_blocks :689-763 and lmult_by_inv_jacob_product :944-981 -- the scripts' default solver,
scripts/utils.py:138-142) against the oracle, standalone and inside constrained leapfrog steps."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.helpers import make_batched, make_fhn_problem

pytestmark = pytest.mark.gpu

VARIANTS = {"noiseless": dict(), "noisy_param": dict(noise=2, sigma=0.2), "gaussian_split": dict(gaussian=True)}


@pytest.fixture(scope="module", params=sorted(VARIANTS), ids=sorted(VARIANTS))
def prob(request):
    return make_fhn_problem(10, 5, 5, n_chains=3, nd=200, **VARIANTS[request.param])


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("part", [0, 1])
def test_newton_projection_matches_oracle(prob, part):
    sysm = prob["system"]
    rng = np.random.default_rng(7)
    q0, xo = prob["q"], prob["xobs"]
    bc = make_batched(prob)
    bc.opts.solver = 1
    bc.set_state(q0, xo, part)
    bc.linearize(True)
    q_in = q0 + 0.02 * rng.standard_normal(q0.shape)
    q_out, status, iters = bc.project_quasi_newton(q_in)
    for i in range(q0.shape[0]):
        pt = sysm.point(q0[i], xo[i], part)
        q_o, mu, it_o, ndq, err = sysm._newton_projection(
            torch.tensor(q_in[i]), torch.tensor(xo[i]), part, pt["jac"], 0.1, 1e-9, 1e-8, 1e10, 50)
        assert status[i] == 0 and err < 1e-9 and ndq < 1e-8
        assert iters[i] == it_o
        assert _rel(q_out[i], q_o.numpy()) < 1e-9
        c = sysm._constr(torch.tensor(q_out[i]), torch.tensor(xo[i]), part).numpy()
        assert np.max(np.abs(c)) < 1e-9
    bc.close()


@pytest.mark.parametrize("part", [0, 1])
def test_newton_leapfrog_steps(prob, part):
    sysm = prob["system"]
    q0, xo = prob["q"], prob["xobs"]
    dt = 0.05
    rng = np.random.default_rng(3)
    p_raw = rng.standard_normal(q0.shape)
    bc = make_batched(prob)
    bc.opts.solver = 1
    bc.set_state(q0, xo, part, p=p_raw)
    bc.linearize(True)
    bc.project_momentum()
    traj = []
    for s in range(3):
        bc.leapfrog_step(dt)
        info = bc.step_info()
        qg, pg, _ = bc.get_state()
        traj.append((qg, pg, info, bc.hamiltonian()))
    for i in range(q0.shape[0]):
        pt = sysm.point(q0[i], xo[i], part)
        p = sysm.project_onto_cotangent_space(torch.tensor(p_raw[i]), pt)
        q = torch.tensor(q0[i])
        for s in range(3):
            q, p, pt, inf = O.leapfrog_step(sysm, q, p, xo[i], part, dt, pt=pt, solver="newton")
            qg, pg, info, hg = traj[s]
            assert info["status"][i] == 0
            assert info["iters_fwd"][i] == inf["n_fwd"] and info["iters_rev"][i] == inf["n_back"]
            assert _rel(qg[i], q.numpy()) < 1e-9
            assert _rel(pg[i], p.numpy()) < 1e-8
            assert abs(hg[i] - sysm.h(q, p, pt)) < 1e-9 * abs(hg[i])
    bc.close()
