"""GPU parity for the variants of the constrained system: noisy observations with a fixed or an
inferred noise scale (mici_extensions.py:353-358, 601-610, 772-791) and the Gaussian splitting
(:1186-1238), against the float64 autodiff oracle.  Same tolerances as test_gpu_parity_small.py."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.helpers import make_batched, make_fhn_problem

pytestmark = pytest.mark.gpu

VARIANTS = {
    "noisy_fixed": dict(noise=1, sigma=0.2),
    "noisy_param": dict(noise=2, sigma=0.2),
    "gaussian_split": dict(gaussian=True),
    "noisy_param_gaussian": dict(noise=2, sigma=0.3, gaussian=True),
}


@pytest.fixture(scope="module", params=sorted(VARIANTS), ids=sorted(VARIANTS))
def prob(request):
    return make_fhn_problem(10, 5, 5, n_chains=3, nd=200, **VARIANTS[request.param])


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("part", [0, 1])
def test_point_quantities(prob, part):
    rng = np.random.default_rng(1)
    q = prob["q"] + 0.05 * rng.standard_normal(prob["q"].shape)   # off the manifold, non-zero noise variables
    bc = make_batched(prob)
    assert bc.dim_q == q.shape[1]
    bc.set_state(q, prob["xobs"], part)
    c = bc.constr()
    bc.linearize(True)
    ld = bc.log_det_sqrt_gram()
    g = bc.grad_log_det_sqrt_gram()
    vct = rng.standard_normal(q.shape)
    nsc = bc.normal_space_component(vct)
    sysm = prob["system"]
    for i in range(q.shape[0]):
        c_o = sysm._constr(torch.tensor(q[i]), torch.tensor(prob["xobs"][i]), part).numpy()
        assert c.shape[1] == c_o.shape[0]
        assert np.max(np.abs(c[i] - c_o)) < 1e-12
        pt = sysm.point(q[i], prob["xobs"][i], part)
        assert abs(ld[i] - pt["ld"]) < 1e-10 * max(1.0, abs(pt["ld"]))
        assert _rel(g[i], pt["grad_ld"].numpy()) < 1e-9
        nsc_o = sysm._normal_space_component(torch.tensor(vct[i]), pt["jac"], pt["chol"]).numpy()
        assert _rel(nsc[i], nsc_o) < 1e-9
    bc.close()


@pytest.mark.parametrize("part", [0, 1])
def test_leapfrog_steps(prob, part):
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
    traj = []
    for s in range(3):
        bc.leapfrog_step(dt)
        info = bc.step_info()
        qg, pg, _ = bc.get_state()
        traj.append((qg, pg, info, bc.hamiltonian()))
    for i in range(n):
        pt = sysm.point(q0[i], xo[i], part)
        p = sysm.project_onto_cotangent_space(torch.tensor(p_raw[i]), pt)
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
