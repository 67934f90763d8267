"""GPU parity for the standard-HMC baseline's target (conditioned_diffusion_neg_log_dens_and_grad,
sde/mici_extensions.py:82-205) and the Adam initialiser built on it
(find_initial_state_by_gradient_descent_noisy_system, :1679-1801) against the float64 autodiff oracle.
Tolerances: value 1e-12 relative, gradient 1e-10 relative (one forward and one reverse sweep; no iteration)."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from oracle.models import fhn, sir

pytestmark = pytest.mark.gpu

CASES = {
    "fhn_fixed": dict(model="fhn", m=fhn, T=8, S=5, dim_u=4, sigma=0.3, noise=1, obs_interval=0.2),
    "fhn_param": dict(model="fhn", m=fhn, T=8, S=5, dim_u=5, sigma=None, noise=2, obs_interval=0.2),
    "sir_fixed": dict(model="sir", m=sir, T=6, S=4, dim_u=4, sigma=1.0, noise=1, obs_interval=1.0),
    "sir_param": dict(model="sir", m=sir, T=6, S=4, dim_u=5, sigma=None, noise=2, obs_interval=1.0),
}


def _problem(cfg, n, seed=3):
    rng = np.random.default_rng(seed)
    m = cfg["m"]
    dim = cfg["dim_u"] + m.dim_v_0 + cfg["T"] * cfg["S"] * m.dim_v
    if cfg["model"] == "fhn":
        y = rng.standard_normal((cfg["T"], 1))
        q = 0.5 * rng.standard_normal((n, dim))
    else:
        y = np.array([3.0, 8.0, 28.0, 75.0, 221.0, 281.0])[: cfg["T"], None]
        q = 0.3 * rng.standard_normal((n, dim))
        q[:, :4] += np.array([-1.0, -0.5, 0.8, 0.0])
    return y, q


def _oracle_funcs(cfg, y, gaussian):
    m = cfg["m"]
    gen_sigma = cfg["sigma"] if cfg["noise"] == 1 else m.generate_σ_y
    return O.conditioned_diffusion_neg_log_dens_and_grad(
        cfg["obs_interval"], cfg["S"], y, cfg["dim_u"], m.dim_v_0, m.dim_v, m.forward_func, m.generate_x_0,
        m.generate_z, gen_sigma, m.obs_func, use_gaussian_splitting=gaussian)


@pytest.mark.parametrize("gaussian", [False, True])
@pytest.mark.parametrize("name", list(CASES))
def test_value_and_gradient(name, gaussian):
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    cfg = CASES[name]
    n = 5
    y, q = _problem(cfg, n)
    bc = BatchedChains(cfg["model"], cfg["obs_interval"], cfg["S"], 3, y, cfg["dim_u"], n, noise=cfg["noise"],
                       sigma_fixed=cfg["sigma"] or 0.0, use_gaussian_splitting=gaussian)
    assert bc.hmc_dim == q.shape[1]
    val, grad, res = bc.neg_log_dens_and_grad(q, gaussian, with_residuals=True)
    nld, vg = _oracle_funcs(cfg, y, gaussian)
    for c in range(n):
        v_o, g_o = vg(q[c])
        assert abs(val[c] - float(v_o)) <= 1e-12 * max(1.0, abs(float(v_o)))
        assert np.max(np.abs(grad[c] - g_o.numpy())) <= 1e-10 * max(1.0, np.max(np.abs(g_o.numpy())))
    # value-only call gives the same numbers
    val2, none = bc.neg_log_dens_and_grad(q, gaussian, with_grad=False)
    assert none is None and np.array_equal(val, val2)
    assert np.isfinite(res).all() and res.shape == (n, cfg["T"])
    bc.close()


def test_reference_surface_and_errors():
    """The drop-in functions: same signature / return conventions as the reference (value; (grad, value)),
    HamiltonianDivergenceError on a non-finite value (:193-204)."""
    from manifold_mcmc_for_diffusions_b200 import mici_extensions as me
    from manifold_mcmc_for_diffusions_b200.example_models import fhn as gfhn
    from manifold_mcmc_for_diffusions_b200.mici_compat.errors import HamiltonianDivergenceError

    cfg = CASES["fhn_param"]
    y, q = _problem(cfg, 1)
    nld, gnld = me.conditioned_diffusion_neg_log_dens_and_grad(
        cfg["obs_interval"], cfg["S"], y, cfg["dim_u"], 2, 2, gfhn.forward_func, gfhn.generate_x_0, gfhn.generate_z,
        gfhn.generate_σ_y, gfhn.obs_func)
    _, vg = _oracle_funcs(cfg, y, False)
    v_o, g_o = vg(q[0])
    assert isinstance(nld(q[0]), float) and abs(nld(q[0]) - float(v_o)) < 1e-12 * abs(float(v_o))
    g, v = gnld(q[0])
    assert abs(v - float(v_o)) < 1e-12 * abs(float(v_o)) and np.max(np.abs(g - g_o.numpy())) < 1e-10 * np.abs(g_o).max()
    bad = q[0].copy()
    bad[cfg["dim_u"] + 2:] = 1e3      # the FHN cubic overflows
    with pytest.raises(HamiltonianDivergenceError):
        nld(bad)


ADAM_CASES = {
    # (case, sigma override, chains, keyword arguments); the last one needs 16 restarts (divergence / slow progress)
    "fhn_fixed": ("fhn_fixed", 0.5, 3, dict(adam_step_size=5e-2, max_iters=150, threshold=2.0, check_iter=10)),
    "sir_param": ("sir_param", None, 3, dict(adam_step_size=5e-2, max_iters=60, threshold=50.0, check_iter=10,
                                             max_num_tries=40)),
    "fhn_restarts": ("fhn_fixed", 0.3, 1, dict(adam_step_size=5e-2, max_iters=60, threshold=2.0, check_iter=10,
                                               max_num_tries=40)),
}


@pytest.mark.parametrize("name", list(ADAM_CASES))
def test_adam_initialiser_matches_oracle(name):
    """Same generator per chain, same Adam constants: the accepted point, its residuals, the number of tries and the
    constraint residual of the resulting state (residuals as noise variables put it on the manifold, :1767-1771)."""
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    case, sigma, n, kw = ADAM_CASES[name]
    cfg = dict(CASES[case])
    if sigma is not None:
        cfg["sigma"] = sigma
    m = cfg["m"]
    y, _ = _problem(cfg, 3)
    if cfg["model"] == "fhn":
        y = 0.5 * y
    bc = BatchedChains(cfg["model"], cfg["obs_interval"], cfg["S"], 3, y, cfg["dim_u"], n, noise=cfg["noise"],
                       sigma_fixed=cfg["sigma"] or 0.0)
    q, tries = bc.init_gradient_descent([np.random.default_rng([7, c]) for c in range(n)], **kw)
    md = {"dim_u": cfg["dim_u"], "dim_v_0": m.dim_v_0, "dim_v": m.dim_v, "num_obs": cfg["T"],
          "num_steps_per_obs": cfg["S"], "δ": cfg["obs_interval"] / cfg["S"], "y_seq": y,
          "generate_z": m.generate_z, "generate_x_0": m.generate_x_0, "forward_func": m.forward_func,
          "obs_func": m.obs_func, "generate_σ": cfg["sigma"] if cfg["noise"] == 1 else m.generate_σ_y}
    for c in range(n):
        u_v, res, t_o, _ = O.find_initial_state_by_gradient_descent_noisy_system(md, np.random.default_rng([7, c]), **kw)
        assert tries[c] == t_o
        assert np.max(np.abs(q[c, : u_v.size] - u_v)) < 1e-8 * max(1.0, np.abs(u_v).max())
        assert np.max(np.abs(q[c, u_v.size:] - res)) < 1e-7 * max(1.0, np.abs(res).max())
    if name == "fhn_restarts":
        assert tries[0] > 1
    assert np.max(np.abs(bc.constr())) < 1e-9 * max(1.0, np.abs(y).max())
    bc.close()
