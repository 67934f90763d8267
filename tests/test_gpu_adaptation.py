"""Per-chain step sizes and the on-device dual-averaging step-size adapter
(mici.adapters.DualAveragingStepSizeAdapter as used at scripts/utils.py:303-306)."""

import numpy as np
import pytest

from tests.helpers import make_batched, make_fhn_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prob():
    return make_fhn_problem(10, 5, 5, n_chains=6, nd=200)


def _run(prob, dts, idx):
    bc = make_batched(prob, n_chains=len(idx))
    rng = np.random.default_rng(3)
    p_raw = rng.standard_normal(prob["q"].shape)
    bc.set_state(prob["q"][idx], prob["xobs"][idx], 0, p=p_raw[idx])
    bc.linearize(True)
    bc.project_momentum()
    if np.ndim(dts):
        bc.set_step_sizes(dts)
        assert np.array_equal(bc.get_step_sizes(), dts)
        bc.leapfrog_step(1.0)       # only the sign of the scalar matters with per-chain step sizes
        bc.leapfrog_step(1.0)
    else:
        bc.leapfrog_step(dts)
        bc.leapfrog_step(dts)
    q, p, _ = bc.get_state()
    info = bc.step_info()
    bc.close()
    return q, p, info


def test_per_chain_step_sizes_match_scalar_runs(prob):
    dts = np.array([0.05, 0.05, 0.05, 0.02, 0.02, 0.02])
    q, p, info = _run(prob, dts, np.arange(6))
    qa, pa, _ = _run(prob, 0.05, np.arange(0, 3))
    qb, pb, _ = _run(prob, 0.02, np.arange(3, 6))
    assert np.all(info["status"] == 0)
    assert np.array_equal(q[:3], qa) and np.array_equal(p[:3], pa)
    assert np.array_equal(q[3:], qb) and np.array_equal(p[3:], pb)


def test_dual_averaging_reaches_target_accept_rate():
    prob = make_fhn_problem(10, 5, 5, n_chains=64, nd=200)
    bc = make_batched(prob)
    bc.set_state(prob["q"], prob["xobs"], 0)
    for it in range(30):                                   # leave the interpolated initial states
        bc.hmc_transition(0.02, 4, 5, it)
    bc.adapt_start(0.02, target=0.8, reg_coefficient=0.1)
    acc = []
    for it in range(30, 330):
        bc.hmc_transition(1.0, 4, 5, it)
        acc.append(bc.transition_stats()["accept_stat"])
    dt = bc.get_step_sizes()
    acc = np.array(acc)
    assert np.all(dt > 0) and np.all(np.isfinite(dt)) and dt.std() > 0
    assert abs(acc[-150:].mean() - 0.8) < 0.08             # adapted towards the target
    assert dt.mean() > 0.03                                 # and away from the conservative start
    bc.adapt_stop(pool=True)
    pooled = bc.get_step_sizes()
    assert np.allclose(pooled, pooled[0]) and pooled[0] > 0
    for it in range(330, 360):
        bc.hmc_transition(1.0, 4, 5, it)
    assert np.array_equal(bc.get_step_sizes(), pooled)      # no adaptation after finalize
    assert np.max(np.abs(bc.constr())) < 1e-8
    bc.close()


def test_init_step_size_search_and_per_chain_adapter_start(prob):
    """Batched DualAveragingStepSizeAdapter._find_and_set_init_step_size: every chain ends on the log-2 boundary of
    |delta H| (one more halving / doubling crosses it), positions are untouched, and the adapter starts from the
    per-chain result."""
    from manifold_mcmc_for_diffusions_b200.adaptation import find_init_step_sizes

    bc = make_batched(prob)
    bc.set_state(prob["q"], prob["xobs"], 0)
    eps, found = find_init_step_sizes(bc, 7, 0)
    assert found.all()
    assert np.all(np.log2(eps) == np.round(np.log2(eps))) and np.all(eps < 1.0)     # halvings of 1.0
    q, _, _ = bc.get_state()
    assert np.array_equal(q, prob["q"])
    assert np.all(bc.step_info()["status"] == 0)
    # re-do the last two trials by hand with the same momenta: eps ends a "too big" search, i.e. |dH(eps)| <= log 2
    # while the step before it (2 eps) failed or changed H by more than log 2
    thr = np.log(2.0)
    dh = {}
    for k, e in (("eps", eps), ("twice", 2.0 * eps)):
        bc.transition_begin(7, 0)
        h0 = bc.hamiltonian()
        bc.set_step_sizes(e)
        bc.transition_steps(1.0, 1)
        st = bc.step_info()["status"]
        with np.errstate(invalid="ignore"):
            d = np.abs(h0 - bc.hamiltonian())
        dh[k] = np.where(st != 0, np.inf, d)
        bc.set_state(prob["q"], prob["xobs"], 0)
    assert np.all(dh["eps"] <= thr)
    assert np.all(~(dh["twice"] <= thr))
    bc.set_step_sizes(None)
    bc.adapt_start(eps, target=0.8, reg_coefficient=0.1)
    assert np.array_equal(bc.get_step_sizes(), eps)
    bc.adapt_stop()
    bc.close()
