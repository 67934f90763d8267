"""Host logic of the batched NUTS transition (nuts.py) against the sequential recursive transition of
mici_compat (the Mici 0.1.10 restatement) on a target both can run on the CPU: an anisotropic Gaussian with a plain
leapfrog integrator and an artificial "projection failure" region.  The batched driver only talks to the chains
object through the mmd_vec_* / transition_* surface, which a NumPy stand-in implements here."""

import numpy as np
import pytest

from manifold_mcmc_for_diffusions_b200.mici_compat.errors import ConvergenceError
from manifold_mcmc_for_diffusions_b200.mici_compat.transitions import MultinomialDynamicIntegrationTransition
from manifold_mcmc_for_diffusions_b200.nuts import BatchedNUTS

SCALES = np.array([1.0, 0.5, 2.0, 0.25, 1.5, 0.8])
FAIL_AT = 2.6       # |q_0| beyond this: the step "fails to converge"


def _grad(q):
    return q / SCALES ** 2


def _leapfrog(q, p, dt):
    p = p - 0.5 * dt * _grad(q)
    q = q + dt * p
    p = p - 0.5 * dt * _grad(q)
    return q, p


def _h(q, p):
    return 0.5 * np.sum(q * q / SCALES ** 2, -1) + 0.5 * np.sum(p * p, -1)


class FakeChains:
    """NumPy stand-in for BatchedChains: the subset of the surface BatchedNUTS drives."""
    VEC_Q, VEC_P = -1, -2

    def __init__(self, q, seed):
        self.q = np.array(q)
        self.n_chains, self.dim = self.q.shape
        self.p = np.zeros_like(self.q)
        self.aux = []
        self.dt = np.zeros(self.n_chains)
        self.status = np.zeros(self.n_chains, dtype=np.int32)
        self.rng = np.random.default_rng(seed)
        self.partition = 0

    def aux_reserve(self, k):
        while len(self.aux) < k:
            self.aux.append(np.zeros_like(self.q))

    def _v(self, i):
        return self.q if i == -1 else self.p if i == -2 else self.aux[i]

    def vec_axpby(self, dst, src, alpha=1.0, beta=0.0, mask=None):
        d, s = self._v(dst), self._v(src)
        new = alpha * s + (beta * d if beta != 0.0 else 0.0)
        if alpha == 0.0:
            new = beta * d if beta != 0.0 else np.zeros_like(d)
        m = slice(None) if mask is None else np.asarray(mask, bool)
        d[m] = new[m]

    def vec_uturn(self, a, d, c, e):
        s = self._v(c) - self._v(d) + self._v(a)
        return np.sum(self._v(a) * s, 1), np.sum(self._v(e) * s, 1)

    def transition_begin(self, seed, it):
        self.p[:] = self.rng.standard_normal(self.q.shape)

    def hamiltonian(self):
        return _h(self.q, self.p)

    def set_step_sizes(self, dt):
        self.dt = np.broadcast_to(np.asarray(dt, float), (self.n_chains,)).copy()

    def set_inactive(self, mask=None, clear_errors=False):
        if clear_errors:
            self.status[:] = 0
        else:
            self.status &= ~16
        if mask is not None:
            self.status[np.asarray(mask, bool)] |= 16

    def relinearize(self):
        pass

    def transition_steps(self, scale, n):
        for c in np.flatnonzero(self.status == 0):
            q, p = _leapfrog(self.q[c], self.p[c], scale * self.dt[c])
            if abs(q[0]) > FAIL_AT:
                self.status[c] |= 1
            else:
                self.q[c], self.p[c] = q, p

    def step_info(self):
        return {"status": self.status.copy()}

    def switch_partition(self):
        self.partition = 1 - self.partition


class _State:
    def __init__(self, pos, mom, dir=1):
        self.pos, self.mom, self.dir = pos, mom, dir

    def copy(self):
        return _State(self.pos.copy(), self.mom.copy(), self.dir)


class _System:
    def h(self, s):
        return _h(s.pos, s.mom)

    def dh_dmom(self, s):
        return s.mom


class _Integrator:
    def __init__(self, step_size):
        self.step_size = step_size

    def step(self, s):
        q, p = _leapfrog(s.pos, s.mom, s.dir * self.step_size)
        if abs(q[0]) > FAIL_AT:
            raise ConvergenceError("failed")
        return _State(q, p, s.dir)


def _run_sequential(eps, n_chain, n_iter, seed, extra):
    rng = np.random.default_rng(seed)
    tr = MultinomialDynamicIntegrationTransition(_System(), _Integrator(eps), max_tree_depth=6,
                                                 do_extra_subtree_checks=extra)
    out = {k: [] for k in ("n_step", "accept_stat", "tree_depth", "convergence_error")}
    qs = []
    for _ in range(n_chain):
        st = _State(rng.standard_normal(len(SCALES)) * SCALES * 0.7, None)
        for _ in range(n_iter):
            st.mom = rng.standard_normal(len(SCALES))
            st, stats = tr.sample(st, rng)
            for k in out:
                out[k].append(stats[k])
            qs.append(st.pos.copy())
    return {k: np.array(v, float) for k, v in out.items()}, np.array(qs)


def _run_batched(eps, n_chain, n_iter, seed, extra):
    rng = np.random.default_rng(seed)
    bc = FakeChains(rng.standard_normal((n_chain, len(SCALES))) * SCALES * 0.7, seed + 1)
    nuts = BatchedNUTS(bc, max_tree_depth=6, do_extra_subtree_checks=extra)
    out = {k: [] for k in ("n_step", "accept_stat", "tree_depth", "convergence_error")}
    qs = []
    for it in range(n_iter):
        stats = nuts.transition(eps, rng, 0, it)
        assert np.all(bc.status == 0)
        for k in out:
            out[k].append(np.asarray(stats[k], float))
        qs.append(bc.q.copy())
    return {k: np.concatenate(v) for k, v in out.items()}, np.concatenate(qs)


def _compare(extra):
    eps = 0.2
    a, qa = _run_sequential(eps, 40, 150, 11, extra)
    b, qb = _run_batched(eps, 400, 15, 12, extra)
    for k in a:
        se = np.hypot(a[k].std() / np.sqrt(len(a[k]) / 4), b[k].std() / np.sqrt(len(b[k]) / 4))
        assert abs(a[k].mean() - b[k].mean()) < 4 * se + 1e-12, (k, a[k].mean(), b[k].mean(), se)
    # tree size distribution, not only its mean
    for d in range(7):
        fa, fb = np.mean(a["tree_depth"] == d), np.mean(b["tree_depth"] == d)
        assert abs(fa - fb) < 0.04, (d, fa, fb)
    # both leave the (truncated) target invariant: second moments of the unconstrained coordinates
    va, vb = qa[:, 1:].var(0), qb[:, 1:].var(0)
    assert np.allclose(va, SCALES[1:] ** 2, rtol=0.25) and np.allclose(vb, SCALES[1:] ** 2, rtol=0.25)
    return a, b


def test_batched_nuts_matches_recursive_transition_with_extra_subtree_checks():
    a, b = _compare(True)
    assert a["convergence_error"].mean() > 0.01        # the failure region is actually visited


def test_batched_nuts_matches_recursive_transition_without_extra_subtree_checks():
    _compare(False)


def test_extra_subtree_checks_shorten_trees():
    a, _ = _run_batched(0.2, 300, 10, 3, True)
    b, _ = _run_batched(0.2, 300, 10, 3, False)
    assert a["n_step"].mean() < b["n_step"].mean()


def test_batched_init_step_size_search_matches_the_sequential_adapter():
    """adaptation.find_init_step_sizes against DualAveragingStepSizeAdapter._find_and_set_init_step_size of the
    Mici restatement, chain by chain, on the same states and momenta."""
    from manifold_mcmc_for_diffusions_b200.adaptation import find_init_step_sizes
    from manifold_mcmc_for_diffusions_b200.mici_compat.adapters import DualAveragingStepSizeAdapter

    rng = np.random.default_rng(5)
    n = 64
    q = rng.standard_normal((n, len(SCALES))) * SCALES * 0.9
    bc = FakeChains(q, 21)
    eps, found = find_init_step_sizes(bc, 0, 0)
    assert found.all() and np.array_equal(bc.q, q) and np.all(bc.status == 0)
    p = bc.p.copy()                               # the momenta the search used
    ad = DualAveragingStepSizeAdapter()
    for c in range(n):
        integ = _Integrator(None)
        ref = ad._find_and_set_init_step_size(_State(q[c].copy(), p[c].copy()), _System(), integ)
        assert ref == eps[c], (c, ref, eps[c])
    assert len(set(eps)) > 1                      # not a trivial case: chains end at different step sizes


class _RecordingRng:
    """Generator whose vector draws are kept: column c is the uniform sequence chain c consumes."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.draws = []

    def random(self, n):
        u = self.rng.random(n)
        self.draws.append(u)
        return u


class _ReplayRng:
    def __init__(self, seq):
        self.seq, self.i = list(seq), 0

    def uniform(self):
        u = self.seq[self.i]
        self.i += 1
        return u


@pytest.mark.parametrize("extra", [True, False])
def test_batched_nuts_equals_recursive_transition_chain_by_chain(extra):
    """Same start, same momenta, same uniforms (direction / sub-tree merges / acceptance in the recursion's draw
    order): n_step, tree_depth, termination flags, accept_stat and the selected proposal must be IDENTICAL per chain.
    An index slip in the check-point bookkeeping of nuts.py (popcount / trailing-ones levels, extra sub-tree checks,
    proposal stack) cannot pass this."""
    n, depth = 96, 6
    rng0 = np.random.default_rng(77)
    q0 = rng0.standard_normal((n, len(SCALES))) * SCALES * 0.9
    bc = FakeChains(q0, 5)
    nuts = BatchedNUTS(bc, max_tree_depth=depth, do_extra_subtree_checks=extra)
    rec = _RecordingRng(123)
    trace = {}
    # FakeChains has no get_state: capture the refreshed momenta through the hook the transition calls
    bc.get_state = lambda: (bc.q.copy(), bc.p.copy(), None)
    stats = nuts.transition(0.23, rec, 0, 0, trace=trace)
    qs, ps, _ = trace["state0"]
    draws = np.array(rec.draws)                   # [n_draws, n]
    tr = MultinomialDynamicIntegrationTransition(_System(), _Integrator(0.23), max_tree_depth=depth,
                                                 do_extra_subtree_checks=extra)
    seen_depths = set()
    for c in range(n):
        st, ref = tr.sample(_State(qs[c].copy(), ps[c].copy()), _ReplayRng(draws[:, c]))
        assert ref["n_step"] == stats["n_step"][c], c
        assert ref["tree_depth"] == stats["tree_depth"][c], c
        assert ref["convergence_error"] == bool(stats["convergence_error"][c]), c
        assert ref["diverging"] == bool(stats["diverging"][c]), c
        assert abs(ref["accept_stat"] - stats["accept_stat"][c]) < 1e-14, c
        assert np.array_equal(st.pos, bc.q[c]), c          # the same leaf was proposed and accepted
        seen_depths.add(ref["tree_depth"])
    assert len(seen_depths) >= 3 and stats["convergence_error"].any() and (stats["n_step"] >= 15).any()
