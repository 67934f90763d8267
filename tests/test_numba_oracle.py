"""The compiled CPU restatement (oracle/numba_chmc.py, the `cpu_baseline` of bench.py) against the autodiff oracle
(oracle/torch_oracle.py): the two are written independently (scalar loops over the block structure vs
torch.func.jacrev / grad over the dense restatement of sde/mici_extensions.py)."""

import warnings

import numpy as np
import pytest
import torch

from oracle import numba_chmc as N
from oracle import torch_oracle as O
from tests.helpers import OBS_INTERVAL, make_fhn_problem

warnings.filterwarnings("ignore", category=Warning, module="numba")


@pytest.mark.parametrize("T,S,R,gaussian", [(10, 5, 5, False), (12, 4, 5, False), (10, 5, 5, True)])
def test_numba_port_matches_autodiff_oracle(T, S, R, gaussian):
    pr = make_fhn_problem(T, S, R, n_chains=1, nd=50, gaussian=gaussian)
    sysm, q, xo = pr["system"], pr["q"][0], pr["xobs"][0]
    y = np.asarray(pr["y"]).reshape(T)
    dl = OBS_INTERVAL / S
    for part in (0, 1):
        bo, bn = N.partition_layout(T, R)[part]
        c = N.constr(q, xo, y, bo, bn, S, dl)
        np.testing.assert_allclose(c, sysm._constr(torch.tensor(q), torch.tensor(xo), part).numpy(), atol=1e-13)
        lin = N.linearize(q, xo, y, bo, bn, S, dl)
        g = N.grad_log_det(q, lin, bo, bn, S, dl)
        pt = sysm.point(q, xo, part)
        assert abs(lin[12] - float(pt["ld"])) < 1e-11
        go = pt["grad_ld"].numpy()
        assert np.abs(g - go).max() <= 1e-10 * np.abs(go).max()
        p0 = np.random.default_rng(part).standard_normal(q.shape[0])
        pp = N.project_momentum(p0, lin, bo, bn, S, N.n_rows(bo, bn))
        po = sysm.project_onto_cotangent_space(torch.tensor(p0), pt).numpy()
        np.testing.assert_allclose(pp, po, atol=1e-11)
        # one full constrained leapfrog step: positions / momenta to 1e-9 relative, identical iteration counts
        qo, pn, _, info = O.leapfrog_step(sysm, q, po, xo, part, 0.05, pt=pt)
        st, q2, p2, _, _, nf, nb = N.leapfrog_step(q, pp, xo, y, lin, g, bo, bn, S, dl, 0.05, gaussian, 1e-9, 1e-8, 1e10,
                                                   50, 2e-8)
        assert st == 0 and (nf, nb) == (info["n_fwd"], info["n_back"])
        assert np.abs(q2 - qo.numpy()).max() <= 1e-9 * np.abs(qo.numpy()).max()
        assert np.abs(p2 - pn.numpy()).max() <= 1e-9 * max(1.0, np.abs(pn.numpy()).max())


def test_numba_init_and_xobs_match_oracle():
    T, S, R = 10, 5, 5
    pr = make_fhn_problem(T, S, R, n_chains=1, nd=50)
    y = np.asarray(pr["y"])

    def gen_init(rng_):
        return np.concatenate((y, rng_.standard_normal(y.shape) * 0.5), -1)

    u, v0 = np.array([0.1, -0.2, 0.3, 0.4]), np.array([0.5, -0.5])
    qo, xo = O.find_initial_state_by_linear_interpolation(pr["system"], np.random.default_rng(3), gen_init, u=u, v_0=v0)
    qn, xn = N.linear_interpolation_init(T, S, y, OBS_INTERVAL, np.random.default_rng(3), u=u, v_0=v0)
    np.testing.assert_allclose(xn, xo.numpy(), atol=0)
    np.testing.assert_allclose(qn, qo.numpy(), rtol=1e-9, atol=1e-9)
    xs = N.generate_x_obs_seq(qn, T, S, OBS_INTERVAL / S)
    np.testing.assert_allclose(xs, xn, atol=1e-9)   # the interpolated path reproduces its targets


def test_numba_hmc_transition_moves_and_stays_on_manifold():
    T, S, R = 10, 5, 5
    pr = make_fhn_problem(T, S, R, n_chains=1, nd=50)
    ch = N.NumbaChain(T, S, R, pr["y"], OBS_INTERVAL)
    ch.set_state(pr["q"][0], pr["xobs"][0], 0)
    rng = np.random.default_rng(5)
    q0 = ch.q.copy()
    acc = sum(ch.hmc_transition(0.05, 4, rng)[0] for _ in range(6))
    assert acc >= 3 and np.abs(ch.q - q0).max() > 1e-3
    bo, bn = ch.parts[ch.partition]
    assert np.abs(N.constr(ch.q, ch.xobs, ch.y, bo, bn, S, ch.dl)).max() < 1e-8
