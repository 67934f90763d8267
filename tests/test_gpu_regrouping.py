"""Chain regrouping (mmd_set_chain_regrouping): chains change CTA tiles at every partition switch, but every chain's
trajectory, accept decision and statistics are bit-identical to a run without regrouping."""

import numpy as np
import pytest

from tests.helpers import make_batched, make_fhn_problem

pytestmark = pytest.mark.gpu


def _run(prob, n, regroup, n_tr=6):
    reps = (n + prob["q"].shape[0] - 1) // prob["q"].shape[0]
    rng = np.random.default_rng(3)
    q = np.tile(prob["q"], (reps, 1))[:n]
    x = np.tile(prob["xobs"], (reps, 1, 1))[:n]
    bc = make_batched(prob, n_chains=n)
    bc.set_chain_offset(1000)
    if regroup:
        bc.set_chain_regrouping(True)
    bc.set_state(q, x, 0)
    stats = []
    for it in range(n_tr):
        # step sizes large enough that iteration counts differ and some chains fail / are rejected
        bc.hmc_transition(0.08 if it % 2 else 0.03, 3, 11, it)
        st = bc.transition_stats()
        sc = bc.slot_chains()
        acc = np.empty(n, dtype=st["accepted"].dtype)
        acc[sc] = st["accepted"]
        aps = np.empty(n)
        aps[sc] = st["accept_stat"]
        stats.append((acc, aps))
    qs, _, xs = bc.get_state()
    sc = bc.slot_chains()
    qo, xo = np.empty_like(qs), np.empty_like(xs)
    qo[sc], xo[sc] = qs, xs
    info = bc.step_info()
    moved = not np.array_equal(sc, np.arange(n))
    part = bc.partition
    c = np.max(np.abs(bc.constr()))
    bc.close()
    return qo, xo, stats, moved, part, c, info


def test_regrouped_chains_reproduce_the_unregrouped_run():
    prob = make_fhn_problem(12, 4, 5, n_chains=6, nd=200)
    n = 40                                      # 5 tiles of 8: chains migrate between tiles
    qa, xa, sa, moved_a, part_a, ca, _ = _run(prob, n, False)
    qb, xb, sb, moved_b, part_b, cb, _ = _run(prob, n, True)
    assert not moved_a and moved_b              # the regrouped run really permuted its slots
    assert part_a == part_b and ca < 1e-7 and cb < 1e-7
    for (acc_a, ap_a), (acc_b, ap_b) in zip(sa, sb):
        assert np.array_equal(acc_a, acc_b)
        assert np.array_equal(ap_a, ap_b)
    assert np.array_equal(qa, qb)               # bit-identical positions, chain by chain
    assert np.array_equal(xa, xb)


def test_regrouping_is_refused_with_per_chain_step_sizes():
    prob = make_fhn_problem(10, 5, 5, n_chains=4, nd=200)
    bc = make_batched(prob)
    bc.set_chain_regrouping(True)
    with pytest.raises(RuntimeError):
        bc.set_step_sizes(np.full(4, 0.05))
    with pytest.raises(RuntimeError):
        bc.adapt_start(0.05)
    bc.set_state(prob["q"], prob["xobs"], 0)     # back to chain order
    assert np.array_equal(bc.slot_chains(), np.arange(4))
    bc.close()
