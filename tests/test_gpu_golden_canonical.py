"""GPU parity at BASELINE.json's canonical size (FHN T=100, S=25, R=5) against committed
oracle-frozen golden vectors (tests/golden/make_golden.py), plus size-independent properties on a
larger batch (constraint satisfaction, tangency, reversibility, energy behaviour)."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "fhn_T100_S25_R5_golden.npz"), allow_pickle=True)


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _chains(gold, part):
    return [c for c in range(gold["q0"].shape[0]) if c % 2 == part]


@pytest.mark.parametrize("part", [0, 1])
def test_canonical_golden(gold, part):
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    idx = _chains(gold, part)
    T, S, R, dt = int(gold["T"]), int(gold["S"]), int(gold["R"]), float(gold["dt"])
    bc = BatchedChains("fhn", 0.2, S, R, gold["y"], 4, len(idx))
    q0, xo, p_raw = gold["q0"][idx], gold["xobs"][idx], gold["p_raw"][idx]
    # constraint at an off-manifold point
    for j, c in enumerate(idx):
        rng = np.random.default_rng([20200710, c])
        # replay the generator to the same draw as make_golden.py
        rng.standard_normal(4); rng.standard_normal(2); rng.standard_normal((T, 1)); rng.standard_normal(q0.shape[1])
        q_off = q0[j] + 0.01 * rng.standard_normal(q0.shape[1])
        bc1 = BatchedChains("fhn", 0.2, S, R, gold["y"], 4, 1)
        bc1.set_state(q_off[None], xo[j][None], part)
        assert np.max(np.abs(bc1.constr()[0] - gold[f"c_off_{c}"])) < 1e-11
        bc1.close()
    bc.set_state(q0, xo, part, p=p_raw)
    bc.linearize(True)
    assert np.max(np.abs(bc.log_det_sqrt_gram() - gold["ld"][idx])) < 1e-9 * np.max(np.abs(gold["ld"][idx]))
    assert _rel(bc.grad_log_det_sqrt_gram(), gold["grad_ld"][idx]) < 1e-9
    assert _rel(bc.normal_space_component(p_raw), gold["nsc"][idx]) < 1e-9
    bc.project_momentum()
    h0 = bc.hamiltonian()
    assert np.max(np.abs(h0 - gold["traj_h"][idx, 0]) / np.abs(h0)) < 1e-9
    for s in range(gold["traj_q"].shape[1]):
        bc.leapfrog_step(dt)
        info = bc.step_info()
        q, p, _ = bc.get_state()
        assert np.all(info["status"] == 0)
        assert np.array_equal(info["iters_fwd"], gold["traj_it"][idx, s, 0])
        assert np.array_equal(info["iters_rev"], gold["traj_it"][idx, s, 1])
        assert _rel(q, gold["traj_q"][idx, s]) < 1e-9        # north-star: 1e-9 relative on positions
        assert _rel(p, gold["traj_p"][idx, s]) < 1e-8
        h = bc.hamiltonian()
        assert np.max(np.abs(h - gold["traj_h"][idx, s + 1]) / np.abs(h)) < 1e-9
    bc.close()


def test_properties_full_size_batch(gold):
    """512 chains at the canonical size: after a leapfrog step every successful chain satisfies the
    constraint to the solver tolerance, keeps its momentum tangent, passes the reverse check, and a
    step back with -dt returns to the start (time reversibility)."""
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    n = 512
    T, S, R = int(gold["T"]), int(gold["S"]), int(gold["R"])
    y = gold["y"]
    rng = np.random.default_rng(99)
    bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
    u = 0.5 * rng.standard_normal((n, 4))
    v0 = rng.standard_normal((n, 2))
    xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
    bc.init_linear_interpolation(u, v0, xo, 0)
    assert np.max(np.abs(bc.constr())) < 1e-8
    bc.linearize(True)
    bc.sample_momentum(7, 0)
    qa, pa, _ = bc.get_state()
    ha = bc.hamiltonian()
    dt = 0.02
    bc.leapfrog_step(dt)
    info = bc.step_info()
    ok = info["status"] == 0
    assert ok.mean() > 0.95
    qb, pb, _ = bc.get_state()
    hb = bc.hamiltonian()
    assert np.max(np.abs(bc.constr()[ok])) < 1e-9          # constraint_tol
    assert np.max(info["rev_dist"][ok]) < 2e-8             # reverse_check_tol
    nsc = bc.normal_space_component(pb)
    assert np.max(np.abs(nsc[ok])) < 1e-7 * np.max(np.abs(pb[ok]))
    assert np.median(np.abs(hb - ha)[ok] / np.abs(ha[ok])) < 1e-3
    assert np.array_equal(qa[~ok], qb[~ok])                # failed chains did not move
    bc.leapfrog_step(-dt)
    info2 = bc.step_info()
    qc, pc, _ = bc.get_state()
    both = ok & (info2["status"] == 0)
    assert np.max(np.abs(qc[both] - qa[both])) < 1e-6 * np.max(np.abs(qa[both]))
    bc.close()


def test_philox_momentum_is_tangent_and_standard(gold):
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    idx = _chains(gold, 0)
    bc = BatchedChains("fhn", 0.2, int(gold["S"]), int(gold["R"]), gold["y"], 4, len(idx))
    bc.set_state(gold["q0"][idx], gold["xobs"][idx], 0)
    bc.linearize(True)
    bc.sample_momentum(123, 5)
    _, p1, _ = bc.get_state()
    bc.sample_momentum(123, 5)
    _, p2, _ = bc.get_state()
    assert np.array_equal(p1, p2)                          # counter-based: reproducible
    bc.sample_momentum(123, 6)
    _, p3, _ = bc.get_state()
    assert not np.array_equal(p1, p3)
    assert abs(p1.std() - 1.0) < 0.05 and abs(p1.mean()) < 0.05
    assert np.max(np.abs(bc.normal_space_component(p1))) < 1e-9
    bc.close()


@pytest.mark.parametrize("part", [0, 1])
@pytest.mark.parametrize("via_switch", [False, True])
def test_full_tiles_agree_with_single_chain(gold, part, via_switch):
    """Copies of the golden chains filling several CTA tiles (every lane of a tile active, blocks of
    unequal length in partition 1) must reproduce the single-chain results bit for bit, also when the
    partition was reached through SwitchPartitionTransition (re-tiling on device)."""
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    idx = _chains(gold, part)
    rep = 20
    q0 = np.tile(gold["q0"][idx], (rep, 1))
    xo = np.tile(gold["xobs"][idx], (rep, 1, 1))
    p = np.tile(gold["p_raw"][idx], (rep, 1))
    m = len(idx)
    bc = BatchedChains("fhn", 0.2, int(gold["S"]), int(gold["R"]), gold["y"], 4, rep * m)
    if via_switch:
        bc.set_state(q0, xo, 1 - part)
        bc.switch_partition()
        bc.set_momentum(p)
    else:
        bc.set_state(q0, xo, part, p=p)
    bc.linearize(True)
    bc.project_momentum()
    dt = float(gold["dt"])
    for s in range(2):
        bc.leapfrog_step(dt)
        info = bc.step_info()
        q, pp, _ = bc.get_state()
        assert np.all(info["status"] == 0)
        assert np.array_equal(info["iters_fwd"], np.tile(gold["traj_it"][idx, s, 0], rep))
        assert np.array_equal(info["iters_rev"], np.tile(gold["traj_it"][idx, s, 1], rep))
        qr = q.reshape(rep, m, -1)
        assert np.array_equal(qr, np.broadcast_to(qr[0], qr.shape))
        # x_obs_seq regenerated on device by the switch differs from the stored one by rounding only
        assert _rel(qr[0], gold["traj_q"][idx, s]) < (1e-7 if via_switch else 1e-9)
    bc.close()
