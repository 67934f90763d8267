"""Shared problem builders for the parity tests (oracle side + seeded inputs)."""

import numpy as np
import torch

from oracle import torch_oracle as O
from oracle.models import fhn, fhn_simulate_y_seq_numpy

Z_TRUE = np.array([0.3, 0.1, 1.5, 0.8])  # fhn_model_noiseless_obs_chmc_experiment.py:41-46
X0_TRUE = np.array([-0.5, 0.2])
OBS_INTERVAL = 0.2
SEED = 20200710


def torch_generators(gen_params):
    """Oracle-side generate_z / generate_x_0 for the FHN run-time generator parameters (include/mmd_b200.h:
    z_i = a_i u_i + b_i, exponentiated where m_i; x_0 = v_0 + c + E z)."""
    gp = np.asarray(gen_params, dtype=np.float64)
    a, b, m = torch.tensor(gp[0:4]), torch.tensor(gp[4:8]), torch.tensor(gp[8:12] != 0)
    c, E = torch.tensor(gp[12:14]), torch.tensor(gp[14:22].reshape(2, 4))

    def generate_z(u):
        lin = a * u[..., :4] + b
        return torch.where(m, torch.exp(lin), lin)

    def generate_x_0(z, v_0):
        return v_0 + c + z @ E.T

    return generate_z, generate_x_0


def make_fhn_problem(T, S, R, n_chains, nd=1000, seed=SEED, noise=0, sigma=0.1, gaussian=False, gen_params=None):
    """Simulated data + oracle system + linear-interpolation initial states for `n_chains` chains
    (restates fhn_model_noiseless_obs_chmc_experiment.py:84-134 with `nd` fine steps per obs).
    noise: 0 noiseless, 1 fixed observation noise scale `sigma`, 2 inferred scale sigma = exp(u[4])
    (fhn_model_noisy_obs_chmc_experiment.py); gaussian: use_gaussian_splitting."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((T * nd, 2))
    y = fhn_simulate_y_seq_numpy(Z_TRUE, X0_TRUE, v, OBS_INTERVAL / nd, nd)
    if noise:
        y = y + sigma * rng.standard_normal(y.shape)
    dim_u = 5 if noise == 2 else 4
    gen_sigma = None if noise == 0 else (float(sigma) if noise == 1 else fhn.generate_σ_y)
    gen_z, gen_x0 = (fhn.generate_z, fhn.generate_x_0) if gen_params is None else torch_generators(gen_params)
    system = O.OracleSystem(
        OBS_INTERVAL, S, R, y, dim_u, 2, 2, fhn.forward_func, gen_x0, gen_z, fhn.obs_func,
        gen_sigma, gaussian, dim_v_0=2,
    )

    def gen_init(rng_):
        return np.concatenate((y, rng_.standard_normal(y.shape) * 0.5), -1)

    qs, xs = [], []
    for c in range(n_chains):
        crng = np.random.default_rng([seed, c])
        u = crng.standard_normal(dim_u) * 0.5
        if noise == 2:
            u[4] = np.log(sigma) + 0.3 * u[4]
        v0 = crng.standard_normal(2)
        q, xo = O.find_initial_state_by_linear_interpolation(system, crng, gen_init, u=u, v_0=v0)
        qs.append(q.numpy())
        xs.append(xo.numpy())
    return dict(T=T, S=S, R=R, y=y, system=system, q=np.stack(qs), xobs=np.stack(xs), noise=noise, sigma=sigma,
                gaussian=gaussian, dim_u=dim_u, gen_params=gen_params)


def make_batched(prob, n_chains=None):
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    n = prob["q"].shape[0] if n_chains is None else n_chains
    return BatchedChains("fhn", OBS_INTERVAL, prob["S"], prob["R"], prob["y"], prob.get("dim_u", 4), n,
                         noise=prob.get("noise", 0), sigma_fixed=prob.get("sigma", 0.0),
                         use_gaussian_splitting=prob.get("gaussian", False), generator_params=prob.get("gen_params"))


def oracle_momentum(prob, q, xobs, part, seed):
    """Projected momentum for one chain from a seeded normal draw (both sides get the same draw)."""
    rng = np.random.default_rng(seed)
    p = torch.tensor(rng.standard_normal(q.shape[0]))
    pt = prob["system"].point(q, xobs, part)
    return prob["system"].project_onto_cotangent_space(p, pt).numpy(), pt
