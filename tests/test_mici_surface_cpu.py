"""CPU tests of the reference-facing host layer: the Mici stand-in (state cache semantics, NUTS and
dual averaging on a toy Gaussian system), the NumPy model mirrors against the oracle, and the
refusal of untagged model callables (no CPU fallback)."""

import numpy as np
import pytest
import torch

from manifold_mcmc_for_diffusions_b200 import example_models, install_reference_aliases, mici_compat, mici_extensions
from manifold_mcmc_for_diffusions_b200.mici_compat.states import ChainState, cache_in_state, cache_in_state_with_aux
from oracle.models import fhn as ofhn


def test_fhn_numpy_model_matches_oracle():
    rng = np.random.default_rng(0)
    m = example_models.fhn
    for _ in range(5):
        u = rng.standard_normal(5)
        z = m.generate_z(u)
        assert np.allclose(z, ofhn.generate_z(torch.tensor(u)).numpy(), rtol=1e-15)
        x, v = rng.standard_normal(2), rng.standard_normal(2)
        f = m.forward_func(z, x, v, 0.008)
        assert np.max(np.abs(f - ofhn.forward_func(torch.tensor(z), torch.tensor(x), torch.tensor(v), 0.008).numpy())) < 1e-14
        v0 = rng.standard_normal(2)
        assert np.allclose(m.generate_x_0(z, v0), ofhn.generate_x_0(torch.tensor(z), torch.tensor(v0)).numpy())
        assert np.isclose(m.generate_σ_y(u), np.exp(u[4]))
    vs = rng.standard_normal((20, 2))
    z = m.generate_z(np.array([-1.2, -2.3, 0.4, 0.8]))
    ys = m.generate_y_seq(z, np.array([-0.5, 0.2]), vs, 0.04, 5)
    yo = ofhn.generate_y_seq(torch.tensor(z), torch.tensor([-0.5, 0.2], dtype=torch.float64), torch.tensor(vs), 0.04, 5).numpy()
    assert ys.shape == (4, 1) and np.max(np.abs(ys - yo)) < 1e-13


def test_untagged_callables_are_rejected():
    m = example_models.fhn
    y = np.zeros((10, 1))
    with pytest.raises(NotImplementedError):
        mici_extensions.ConditionedDiffusionConstrainedSystem(
            0.2, 5, 5, y, 4, 2, 2, lambda z, x, v, d: x, m.generate_x_0, m.generate_z, m.obs_func)
    with pytest.raises(ValueError):
        mici_extensions.ConditionedDiffusionConstrainedSystem(
            0.2, 5, 5, y, 4, 2, 2, m.forward_func, m.generate_x_0, m.generate_z, m.obs_func,
            use_gaussian_splitting=True, metric=mici_compat.matrices.IdentityMatrix())


def test_split_helpers():
    a = np.arange(12.0)
    parts = mici_extensions.split_and_reshape(a, ((2,), (2, 3), (4,)))
    assert parts[0].shape == (2,) and parts[1].shape == (2, 3) and parts[2].shape == (4,)
    assert np.array_equal(np.concatenate([p.ravel() for p in parts]), a)


def test_state_cache_semantics():
    class Sys:
        calls = 0

        @cache_in_state("pos")
        def f(self, state):
            Sys.calls += 1
            return float(np.sum(state.pos))

        @cache_in_state("pos")
        def g(self, state):
            return 2 * float(np.sum(state.pos))

        @cache_in_state_with_aux("pos", "g")
        def h(self, state):
            return 3.0, 7.0

    s = Sys()
    st = ChainState(pos=np.ones(3), mom=None, dir=1, _call_counts={})
    assert s.f(st) == 3.0 and s.f(st) == 3.0 and Sys.calls == 1
    cp = st.copy()
    assert s.f(cp) == 3.0 and Sys.calls == 1            # cache travels with the copy
    cp.pos = np.zeros(3)
    assert s.f(cp) == 0.0 and Sys.calls == 2            # assignment invalidates
    assert s.f(st) == 3.0 and Sys.calls == 2            # the original is untouched
    assert s.h(st) == 3.0 and s.g(st) == 7.0            # aux output cached under g's key
    assert st._call_counts is cp._call_counts           # shared call counts


class _GaussSystem(mici_compat.systems.System):
    def __init__(self):
        super().__init__(lambda q: 0.5 * float(q @ q), lambda q: (q, 0.5 * float(q @ q)))

    def h1(self, state):
        return self.neg_log_dens(state)

    def dh1_dpos(self, state):
        return self.grad_neg_log_dens(state)

    def h2(self, state):
        return 0.5 * float(state.mom @ state.mom)

    def dh2_dmom(self, state):
        return state.mom

    def h2_flow(self, state, dt):
        state.pos = state.pos + dt * state.mom

    def sample_momentum(self, state, rng):
        return rng.standard_normal(state.pos.shape)


def test_nuts_and_dual_averaging_on_gaussian():
    system = _GaussSystem()
    integrator = mici_compat.integrators.LeapfrogIntegrator(system)
    rng = np.random.default_rng(1)
    sampler = mici_compat.samplers.MarkovChainMonteCarloMethod(rng, {
        "momentum": mici_compat.transitions.IndependentMomentumTransition(system),
        "integration": mici_compat.transitions.MultinomialDynamicIntegrationTransition(system, integrator),
    })
    adapter = mici_compat.adapters.DualAveragingStepSizeAdapter(0.8, log_step_size_reg_coefficient=0.1)
    init = [ChainState(pos=rng.standard_normal(5), mom=rng.standard_normal(5), dir=1, _call_counts={}) for _ in range(2)]
    states, traces, stats = sampler.sample_chains_with_adaptive_warm_up(
        150, 600, init, trace_funcs=[lambda s: {"pos": s.pos}], adapters={"integration": [adapter]})
    pos = np.concatenate(traces["pos"])
    assert abs(pos.mean()) < 0.15 and abs(pos.var() - 1.0) < 0.2
    acc = np.mean(np.concatenate(stats["integration"]["accept_stat"]))
    assert 0.6 < acc < 0.98
    assert 0.2 < integrator.step_size < 3.0


def test_reference_aliases():
    import sys

    sde = install_reference_aliases(force_mici_compat=True)
    import mici
    import sde.mici_extensions as me  # noqa: F401

    assert mici.integrators.ConstrainedLeapfrogIntegrator is mici_compat.integrators.ConstrainedLeapfrogIntegrator
    assert sde.example_models.fhn.dim_x == 2
    for k in [k for k in sys.modules if k == "mici" or k.startswith("mici.") or k == "sde" or k.startswith("sde.")]:
        del sys.modules[k]


def test_sir_numpy_model_matches_oracle():
    from oracle.models import sir as osir

    rng = np.random.default_rng(4)
    m = example_models.sir
    for _ in range(5):
        u = 0.3 * rng.standard_normal(5)
        z = m.generate_z(u)
        assert np.allclose(z, osir.generate_z(torch.tensor(u)).numpy(), rtol=1e-15)
        x = np.array([6.5, 2.0, 0.5]) + 0.2 * rng.standard_normal(3)
        v = rng.standard_normal(3)
        ref = osir.forward_func(torch.tensor(z), torch.tensor(x), torch.tensor(v), 0.05).numpy()
        assert np.max(np.abs(m.forward_func(z, x, v, 0.05) - ref)) < 1e-13
        assert np.allclose(m.generate_x_0(z, np.array([0.3])), osir.generate_x_0(torch.tensor(z), torch.tensor([0.3], dtype=torch.float64)).numpy())
    xc = np.array([-600.0, 1.0, 0.2])
    ref = osir.forward_func(torch.tensor(z), torch.tensor(xc), torch.tensor(v), 0.05).numpy()
    assert np.allclose(m.forward_func(z, xc, v, 0.05), ref) and m.forward_func(z, xc, v, 0.05)[0] == -500.0
