"""OnlineBlockDiagonalMetricAdapter (sde/mici_extensions.py:1804-1931): Welford accumulation per chain, parallel
combination across chains, regularisation and the block-diagonal metric, against NumPy's sample covariance."""

import numpy as np
import pytest

from manifold_mcmc_for_diffusions_b200 import mici_extensions as me
from manifold_mcmc_for_diffusions_b200.mici_compat.errors import AdaptationError


class _State:
    def __init__(self, pos):
        self.pos = pos


class _Sys:
    metric = None


class _Tr:
    def __init__(self):
        self.system = _Sys()


def _run(ad, draws):
    st = ad.initialize(_State(draws[0]), None)
    for d in draws:
        ad.update(st, _State(d), None, None)
    return st


def test_single_chain_matches_numpy_covariance():
    rng = np.random.default_rng(0)
    draws = rng.standard_normal((50, 9)) @ rng.standard_normal((9, 9))
    ad = me.OnlineBlockDiagonalMetricAdapter(4, reg_iter_offset=5, reg_scale=1e-3)
    tr = _Tr()
    ad.finalize(_run(ad, draws), tr)
    n = 50
    cov = np.cov(draws[:, :4].T) * n / (5 + n) + 1e-3 * (5 / (5 + n)) * np.eye(4)
    blocks = tr.system.metric.blocks
    assert np.allclose(blocks[0].array, np.linalg.inv(cov), rtol=1e-10)
    assert type(blocks[1]).__name__ == "IdentityMatrix"


def test_multi_chain_combination_equals_pooled_covariance():
    rng = np.random.default_rng(1)
    chains = [rng.standard_normal((k, 7)) + i for i, k in enumerate((20, 35, 11))]
    ad = me.OnlineBlockDiagonalMetricAdapter(3)
    tr = _Tr()
    ad.finalize([_run(ad, c) for c in chains], tr)
    pooled = np.concatenate(chains)[:, :3]
    n = pooled.shape[0]
    cov = np.cov(pooled.T) * n / (5 + n) + 1e-3 * (5 / (5 + n)) * np.eye(3)
    assert np.allclose(tr.system.metric.blocks[0].array, np.linalg.inv(cov), rtol=1e-10)


def test_needs_two_samples():
    ad = me.OnlineBlockDiagonalMetricAdapter(2)
    with pytest.raises(AdaptationError):
        ad.finalize(_run(ad, np.zeros((1, 4))), _Tr())
