"""Run-time generator parameters (priors without recompiling; mmd_set_generator_params): the CUDA path with a random
parameter set and with the notebook's against the oracle built from the same generate_z / generate_x_0, and the
`fhn_notebook` model id against the explicit parameter vector."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.helpers import make_batched, make_fhn_problem
from manifold_mcmc_for_diffusions_b200 import BatchedChains
from manifold_mcmc_for_diffusions_b200.example_models import fhn as M

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _random_params():
    rng = np.random.default_rng(21)
    return M.generator_params(scale=rng.uniform(0.4, 1.2, 4), shift=0.3 * rng.standard_normal(4), exp_mask=(1, 1, 0, 1),
                              x0_shift=0.2 * rng.standard_normal(2), x0_z=0.3 * rng.standard_normal((2, 4)))


@pytest.mark.parametrize("which", ["random", "notebook"])
@pytest.mark.parametrize("noise", [0, 2])
def test_point_and_step_with_custom_generators(which, noise):
    gp = _random_params() if which == "random" else M.NOTEBOOK_GENERATOR_PARAMS
    prob = make_fhn_problem(10, 5, 5, n_chains=3, nd=200, noise=noise, gen_params=gp)
    sysm, q0, xo = prob["system"], prob["q"], prob["xobs"]
    rng = np.random.default_rng(3)
    p_raw = rng.standard_normal(q0.shape)
    for part in (0, 1):
        bc = make_batched(prob)
        assert np.array_equal(bc.get_generator_params(), gp)
        bc.set_state(q0, xo, part, p=p_raw)
        c = bc.constr()
        bc.linearize(True)
        ld, g = bc.log_det_sqrt_gram(), bc.grad_log_det_sqrt_gram()
        bc.project_momentum()
        bc.leapfrog_step(0.05)
        info = bc.step_info()
        qg, pg, _ = bc.get_state()
        for i in range(q0.shape[0]):
            c_o = sysm._constr(torch.tensor(q0[i]), torch.tensor(xo[i]), part).numpy()
            assert np.max(np.abs(c[i] - c_o)) < 1e-12
            pt = sysm.point(q0[i], xo[i], part)
            assert abs(ld[i] - pt["ld"]) < 1e-10 * max(1.0, abs(pt["ld"]))
            assert _rel(g[i], pt["grad_ld"].numpy()) < 1e-9
            p = sysm.project_onto_cotangent_space(torch.tensor(p_raw[i]), pt)
            q, p, pt, inf = O.leapfrog_step(sysm, q0[i], p, xo[i], part, 0.05, pt=pt)
            assert info["status"][i] == 0
            assert info["iters_fwd"][i] == inf["n_fwd"] and info["iters_rev"][i] == inf["n_back"]
            assert _rel(qg[i], q.numpy()) < 1e-9 and _rel(pg[i], p.numpy()) < 1e-8
        bc.close()


def test_notebook_model_id_is_the_notebook_parameter_set():
    prob = make_fhn_problem(10, 5, 5, n_chains=3, nd=200, gen_params=M.NOTEBOOK_GENERATOR_PARAMS)
    out = []
    for kw in (dict(model="fhn_notebook"), dict(model="fhn", generator_params=M.NOTEBOOK_GENERATOR_PARAMS)):
        bc = BatchedChains(kw["model"], 0.2, 5, 5, prob["y"], 4, 3, generator_params=kw.get("generator_params"))
        assert np.array_equal(bc.get_generator_params(), M.NOTEBOOK_GENERATOR_PARAMS)
        bc.set_state(prob["q"], prob["xobs"], 0, p=np.random.default_rng(1).standard_normal(prob["q"].shape))
        bc.linearize(True)
        bc.project_momentum()
        bc.leapfrog_step(0.05)
        out.append(bc.get_state()[:2] + (bc.grad_log_det_sqrt_gram(),))
        bc.close()
    assert all(np.array_equal(a, b) for a, b in zip(*out))


def test_wrong_parameter_count_is_rejected():
    from manifold_mcmc_for_diffusions_b200._lib import MmdError

    prob = make_fhn_problem(10, 5, 5, n_chains=1, nd=100)
    bc = make_batched(prob)
    with pytest.raises(MmdError):
        bc.set_generator_params(np.zeros(7))
    bc.close()
