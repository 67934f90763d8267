"""Oracle model restatements vs a fresh SymPy derivation that follows sde/integrators.py (CPU)."""

import numpy as np
import sympy as sp
import torch

from oracle import models as Mo


def test_fhn_step_matches_sympy_derivation():
    f, sy = Mo.derive_fhn_step(simplify=False)
    args = list(sy["z"]) + list(sy["x"]) + list(sy["v"]) + [sy["delta"]]
    fn = sp.lambdify(args, f, "numpy")
    rng = np.random.default_rng(0)
    for _ in range(20):
        z = np.array([0.3, 0.1, 1.5, 0.8]) * np.exp(0.3 * rng.standard_normal(4))
        x = rng.standard_normal(2)
        v = rng.standard_normal(2)
        dl = 0.008
        ref = np.asarray(fn(*z, *x, *v, dl), dtype=float).ravel()
        got = Mo.fhn_forward_func(torch.tensor(z), torch.tensor(x), torch.tensor(v), dl).numpy()
        assert np.max(np.abs(ref - got)) < 1e-14 * max(1.0, np.max(np.abs(ref)))


def test_sir_step_matches_sympy_derivation():
    f, sy = Mo.derive_sir_step(simplify=False)
    args = list(sy["z"]) + list(sy["x"]) + list(sy["v"]) + [sy["delta"]]
    fn = sp.lambdify(args, f, "numpy")
    rng = np.random.default_rng(1)
    for _ in range(20):
        z = np.array([0.5, 0.3, 0.2, 0.1]) * np.exp(0.2 * rng.standard_normal(4))
        x = np.array([np.log(700.0), np.log(20.0), 0.3]) + 0.2 * rng.standard_normal(3)
        v = rng.standard_normal(3)
        dl = 0.05
        ref = np.asarray(fn(*z, *x, *v, dl), dtype=float).ravel()
        got = Mo.sir_forward_func(torch.tensor(z), torch.tensor(x), torch.tensor(v), dl).numpy()
        assert np.max(np.abs(ref - got)) < 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_sir_clip_semantics():
    # sir.py:54-70: components at or below -500 are frozen
    z = torch.tensor([0.5, 0.3, 0.2, 0.1])
    x = torch.tensor([-600.0, 1.0, 0.1])
    out = Mo.sir_forward_func(z, x, torch.zeros(3), 0.05)
    assert out[0].item() == -500.0 and torch.isfinite(out).all()


def test_numpy_simulator_matches_torch_scan():
    rng = np.random.default_rng(2)
    v = rng.standard_normal((40, 2))
    z = np.array([0.3, 0.1, 1.5, 0.8])
    x0 = np.array([-0.5, 0.2])
    a = Mo.fhn_simulate_y_seq_numpy(z, x0, v, 0.01, 10)
    b = Mo.fhn_generate_y_seq(torch.tensor(z), torch.tensor(x0), torch.tensor(v), 0.01, 10).numpy()
    assert np.max(np.abs(a - b)) < 1e-13
