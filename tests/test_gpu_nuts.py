"""Batched dynamic-HMC (NUTS) transition built from the mmd_vec_* device primitives (nuts.py):
invariants on a small problem (states stay on the manifold of the switched partition, statistics in range,
trees grow beyond one step, error-free chains move), and the vector primitives themselves."""

import numpy as np
import pytest

from tests.helpers import make_batched, make_fhn_problem

pytestmark = pytest.mark.gpu


def test_vector_primitives():
    prob = make_fhn_problem(10, 5, 5, n_chains=12, nd=200)
    bc = make_batched(prob)
    rng = np.random.default_rng(0)
    q0 = prob["q"]
    p0 = rng.standard_normal(q0.shape)
    bc.set_state(q0, prob["xobs"], 1, p=p0)
    bc.aux_reserve(4)
    mask = np.arange(12) % 2 == 0
    bc.vec_axpby(0, bc.VEC_Q)                       # aux0 = q
    bc.vec_axpby(1, bc.VEC_P)                       # aux1 = p
    bc.vec_axpby(2, bc.VEC_P, alpha=2.0)            # aux2 = 2 p
    bc.vec_axpby(2, 0, alpha=1.0, beta=1.0, mask=mask)   # aux2 += q on even chains
    bc.vec_axpby(3, 3, alpha=0.0, beta=0.0)         # aux3 = 0
    d1, d2 = bc.vec_uturn(1, 3, 2, bc.VEC_P)        # s = aux2 - 0 + p ; (p . s, p . s)
    s = 3.0 * p0 + np.where(mask[:, None], q0, 0.0)
    ref = np.sum(p0 * s, axis=1)
    assert np.allclose(d1, ref, rtol=1e-12) and np.allclose(d2, ref, rtol=1e-12)
    bc.vec_axpby(bc.VEC_P, 0, mask=mask)            # p <- q on even chains
    _, p, _ = bc.get_state()
    assert np.array_equal(p[mask], q0[mask]) and np.array_equal(p[~mask], p0[~mask])
    bc.close()


def test_nuts_transitions_small_problem():
    from manifold_mcmc_for_diffusions_b200.nuts import BatchedNUTS

    prob = make_fhn_problem(10, 5, 5, n_chains=24, nd=200)
    bc = make_batched(prob)
    bc.set_state(prob["q"], prob["xobs"], 0)
    for it in range(30):
        bc.hmc_transition(0.02, 4, 3, it)
    nuts = BatchedNUTS(bc, max_tree_depth=5)
    rng = np.random.default_rng(5)
    qa, _, _ = bc.get_state()
    n_steps, acc = [], []
    for it in range(30, 42):
        part = bc.partition
        st = nuts.transition(0.05, rng, 3, it)
        assert bc.partition == (part + 1) % 2
        assert np.all((st["accept_stat"] >= 0) & (st["accept_stat"] <= 1))
        assert np.all(st["n_step"] >= 1) and np.all(st["n_step"] <= 2 ** 5 - 1)
        assert np.all(st["tree_depth"] <= 5)
        assert np.max(np.abs(bc.constr())) < 1e-7           # on the manifold of the new partition
        n_steps.append(st["n_step"].mean())
        acc.append(st["accept_stat"].mean())
    qb, _, _ = bc.get_state()
    assert np.mean(n_steps) > 2.0                            # trees actually double
    assert 0.3 < np.mean(acc) <= 1.0
    assert np.mean(np.any(qa != qb, axis=1)) > 0.9           # the chains move
    info = bc.step_info()
    assert np.all(info["status"] == 0)                       # parked / error bits cleared at the end
    bc.close()
