"""Self-consistency invariants of the oracle: the reference ships no tests or golden
vectors, so these are what pins the restatement.  CPU only, small sizes."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.helpers import make_fhn_problem


@pytest.fixture(scope="module")
def prob():
    return make_fhn_problem(20, 5, 5, n_chains=1, nd=100)


def _dense_jacobian(sysm, q, xobs, part):
    return torch.func.jacrev(lambda q_: sysm._constr(q_, xobs, part))(q)


def test_partition_shapes_match_reference_logic(prob):
    sysm = prob["system"]
    # T=20, R=5: partition 0 = (5, 2x5, 5); partition 1 = (2, 3x5, 3)   (mici_extensions.py:327-351)
    assert sysm.y_subseq_shapes[0] == ((5,), (2, 5), (5,))
    assert sysm.y_subseq_shapes[1] == ((2,), (3, 5), (3,))
    assert sysm.num_partition == 2


@pytest.mark.parametrize("part", [0, 1])
def test_init_satisfies_constraint_and_regenerates_xobs(prob, part):
    sysm = prob["system"]
    q, xo = torch.tensor(prob["q"][0]), torch.tensor(prob["xobs"][0])
    assert float(sysm._constr(q, xo, part).abs().max()) < 1e-13
    assert float((sysm._generate_x_obs_seq(q) - xo).abs().max()) < 1e-13


@pytest.mark.parametrize("part", [0, 1])
def test_block_jacobian_equals_dense_jacobian(prob, part):
    sysm = prob["system"]
    rng = np.random.default_rng(5)
    q = torch.tensor(prob["q"][0] + 0.05 * rng.standard_normal(prob["q"][0].shape))
    xo = torch.tensor(prob["xobs"][0])
    J = _dense_jacobian(sysm, q, xo, part)
    jac = sysm._jacob_constr_blocks(q, xo, part)
    vct = torch.tensor(rng.standard_normal(q.shape[0]))
    lam = torch.tensor(rng.standard_normal(J.shape[0]))
    assert float((sysm._lmult_by_jacob_constr(*jac, vct) - J @ vct).abs().max()) < 1e-11
    assert float((sysm._rmult_by_jacob_constr(*jac, lam) - J.T @ lam).abs().max()) < 1e-11
    chol = sysm._chol_gram_blocks(*jac)
    G = J @ J.T
    # Woodbury solve and log-determinant against the dense Gram matrix (:800-810, :915-942)
    sol = sysm._lmult_by_inv_gram(*jac, *chol, lam)
    assert float((G @ sol - lam).abs().max()) < 1e-8 * float(lam.abs().max())
    ld = float(sysm._log_det_sqrt_gram_from_chol(*chol))
    assert abs(ld - 0.5 * float(torch.linalg.slogdet(G)[1])) < 1e-9 * max(1.0, abs(ld))


def test_grad_log_det_matches_dense_autodiff(prob):
    sysm = prob["system"]
    q, xo = torch.tensor(prob["q"][0]), torch.tensor(prob["xobs"][0])

    def dense_ld(q_):
        J = torch.func.jacrev(lambda qq: sysm._constr(qq, xo, 0))(q_)
        return 0.5 * torch.linalg.slogdet(J @ J.T)[1]

    g_dense = torch.func.grad(dense_ld)(q)
    g = sysm.point(q, xo, 0)["grad_ld"]
    assert float((g - g_dense).abs().max()) < 1e-8 * float(g_dense.abs().max())


def test_leapfrog_step_tangency_reversibility_and_energy(prob):
    sysm = prob["system"]
    q, xo = torch.tensor(prob["q"][0]), torch.tensor(prob["xobs"][0])
    rng = np.random.default_rng(6)
    pt = sysm.point(q, xo, 0)
    p = sysm.project_onto_cotangent_space(torch.tensor(rng.standard_normal(q.shape[0])), pt)
    assert float(sysm._lmult_by_jacob_constr(*pt["jac"], p).abs().max()) < 1e-11
    dt = 0.05
    q1, p1, pt1, info = O.leapfrog_step(sysm, q, p, xo, 0, dt, pt=pt)
    assert float(sysm._constr(q1, xo, 0).abs().max()) < 1e-9
    assert float(sysm._lmult_by_jacob_constr(*pt1["jac"], p1).abs().max()) < 1e-10
    assert info["rev_diff"] < 2e-8
    # time reversal: stepping back with dt -> -dt returns to the start
    q0, p0, _, _ = O.leapfrog_step(sysm, q1, p1, xo, 0, -dt, pt=pt1)
    assert float((q0 - q).abs().max()) < 1e-7 and float((p0 - p).abs().max()) < 1e-6
    # energy error shrinks ~ dt^2
    e1 = abs(sysm.h(q1, p1, pt1) - sysm.h(q, p, pt))
    q2, p2, pt2, _ = O.leapfrog_step(sysm, q, p, xo, 0, dt / 2, pt=pt)
    e2 = abs(sysm.h(q2, p2, pt2) - sysm.h(q, p, pt))
    assert e2 < e1


def test_newton_and_quasi_newton_agree(prob):
    sysm = prob["system"]
    q, xo = torch.tensor(prob["q"][0]), torch.tensor(prob["xobs"][0])
    rng = np.random.default_rng(7)
    pt = sysm.point(q, xo, 0)
    p = sysm.project_onto_cotangent_space(torch.tensor(rng.standard_normal(q.shape[0])), pt)
    qa, pa, _, ia = O.leapfrog_step(sysm, q, p, xo, 0, 0.05, pt=pt, solver="quasi_newton")
    qb, pb, _, ib = O.leapfrog_step(sysm, q, p, xo, 0, 0.05, pt=pt, solver="newton")
    assert ib["n_fwd"] <= ia["n_fwd"]
    assert float((qa - qb).abs().max()) < 1e-7
