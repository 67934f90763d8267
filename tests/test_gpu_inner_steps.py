"""GPU parity of ConstrainedLeapfrogIntegrator.step with n_inner_step > 1 (Mici's inner h2 steps; the reference's
--num-inner-h2-step, scripts/utils.py:132, 286) against the oracle's restatement of Mici's _step_b loop, and the
roll-back of chains that fail in an inner step."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.helpers import make_batched, make_fhn_problem

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def prob():
    return make_fhn_problem(10, 5, 5, n_chains=3, nd=200)


@pytest.mark.parametrize("solver", [0, 1])
@pytest.mark.parametrize("n_inner", [2, 3])
@pytest.mark.parametrize("part", [0, 1])
def test_inner_steps_match_oracle(prob, part, n_inner, solver):
    sysm = prob["system"]
    q0, xo = prob["q"], prob["xobs"]
    dt = 0.09
    rng = np.random.default_rng(11)
    p_raw = rng.standard_normal(q0.shape)
    bc = make_batched(prob)
    bc.opts.solver = solver
    bc.set_state(q0, xo, part, p=p_raw)
    bc.linearize(True)
    bc.project_momentum()
    traj = []
    for s in range(2):
        bc.leapfrog_step(dt, n_inner_step=n_inner)
        qg, pg, _ = bc.get_state()
        traj.append((qg, pg, bc.step_info(), bc.hamiltonian()))
    for i in range(q0.shape[0]):
        pt = sysm.point(q0[i], xo[i], part)
        p = sysm.project_onto_cotangent_space(torch.tensor(p_raw[i]), pt)
        q = torch.tensor(q0[i])
        for s in range(2):
            q, p, pt, inf = O.leapfrog_step(sysm, q, p, xo[i], part, dt, pt=pt, n_inner_step=n_inner,
                                            solver="newton" if solver else "quasi_newton")
            qg, pg, info, hg = traj[s]
            assert info["status"][i] == 0
            # the step info reports the LAST inner step's solves
            assert info["iters_fwd"][i] == inf["n_fwd_inner"][-1] and info["iters_rev"][i] == inf["n_back_inner"][-1]
            assert _rel(qg[i], q.numpy()) < 1e-9
            assert _rel(pg[i], p.numpy()) < 1e-8
            assert abs(hg[i] - sysm.h(q, p, pt)) < 1e-9 * abs(hg[i])
    bc.close()


def test_one_inner_step_is_the_plain_step(prob):
    q0, xo = prob["q"], prob["xobs"]
    rng = np.random.default_rng(12)
    p_raw = rng.standard_normal(q0.shape)
    out = []
    for n_inner in (None, 1):
        bc = make_batched(prob)
        bc.set_state(q0, xo, 0, p=p_raw)
        bc.linearize(True)
        bc.project_momentum()
        if n_inner is None:
            bc.leapfrog_step(0.05)
        else:
            bc.leapfrog_step(0.05, n_inner_step=1)
        out.append(bc.get_state()[:2])
        bc.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_failure_in_an_inner_step_rolls_back():
    """Chains that fail in ANY inner step (the first or, after one or two committed inner steps, a later one) must be left
    bit for bit where they were, with the cached quantities of that point (grad log det, Hamiltonian); the others move.
    Iteration budgets around the typical iteration count make some chains fail early, some late and some not at all."""
    prob = make_fhn_problem(10, 5, 5, n_chains=24, nd=200)
    q0, xo = prob["q"], prob["xobs"]
    rng = np.random.default_rng(13)
    p_raw = rng.standard_normal(q0.shape)
    bc = make_batched(prob)
    n_failed = n_ok = 0
    for max_iters in (1, 4, 5, 6, 7, 50):
        bc.opts.max_iters = 50
        bc.set_state(q0, xo, 0, p=p_raw)
        bc.linearize(True)
        bc.project_momentum()
        bc.leapfrog_step(0.05, n_inner_step=2)          # move first: both state slots now hold other points
        assert (bc.step_info()["status"] == 0).all()
        qb, pb, _ = bc.get_state()
        g1, h1 = bc.grad_log_det_sqrt_gram(), bc.hamiltonian()
        bc.opts.max_iters = max_iters
        bc.leapfrog_step(0.24, n_inner_step=3)
        st = bc.step_info()["status"]
        qc, pc, _ = bc.get_state()
        gc, hc = bc.grad_log_det_sqrt_gram(), bc.hamiltonian()
        bad = st != 0
        n_failed += int(bad.sum()); n_ok += int((~bad).sum())
        assert np.array_equal(qb[bad], qc[bad]) and np.array_equal(pb[bad], pc[bad])
        if bad.any():
            assert _rel(gc[bad], g1[bad]) < 1e-12 and np.max(np.abs(hc[bad] - h1[bad])) < 1e-12 * np.max(np.abs(h1))
        assert all(not np.array_equal(qb[i], qc[i]) for i in np.flatnonzero(~bad))
        # the rolled-back chains carry on from there
        bc.opts.max_iters = 50
        bc.leapfrog_step(0.05, n_inner_step=2)
        assert (bc.step_info()["status"][bad] == 0).all()
    assert n_failed > 0 and n_ok > 0
    bc.close()
