"""Hypoelliptic FitzHugh-Nagumo model: NumPy mirror of ``sde/example_models/fhn.py`` (same names).

``forward_func`` is the strong-order-1.5 step (fhn.py:27-34 <- integrators.py:46-63) in the closed
form the CUDA functor is generated from (tools/gen_model_code.py); here it serves the host-side
helpers the scripts call directly (``generate_y_seq`` for data simulation,
fhn_model_noiseless_obs_chmc_experiment.py:91-93, ``generate_z`` / ``generate_x_0`` in trace functions)."""
import math

import numpy as np

dim_x = 2
dim_w = 1
dim_z = 4
dim_v_0 = dim_x
dim_v = 2 * dim_w
_SQRT3 = math.sqrt(3.0)


def _tag(func):
    func._mmd_model = "fhn"
    return func


@_tag
def forward_func(z, x, v, δ):
    σ, ε, γ, β = z[0], z[1], z[2], z[3]
    x0, x1 = x[..., 0], x[..., 1]
    v0, v1 = v[..., 0], v[..., 1]
    P = x0 ** 3 - x0 + x1
    Q = β + γ * x0 - x1
    a0 = -P / ε
    dζ = δ ** 1.5 * (v0 + v1 / _SQRT3) / 2
    f0 = x0 + δ * a0 + (δ ** 2 / 2) * (((1 - 3 * x0 ** 2) / ε) * a0 - Q / ε) - (σ / ε) * dζ
    f1 = x1 + δ * Q + σ * math.sqrt(δ) * v0 + (δ ** 2 / 2) * (γ * a0 - Q) - σ * dζ
    return np.stack([f0, f1], -1)


@_tag
def obs_func(x_seq):
    return x_seq[..., 0:1]


@_tag
def generate_z(u):
    # [σ, ϵ, γ, β]
    return np.stack([np.exp(u[..., 0]), np.exp(u[..., 1]), np.exp(u[..., 2]), u[..., 3]], -1)


@_tag
def generate_σ_y(u):
    return np.exp(u[..., dim_z])


@_tag
def generate_x_0(z, v_0):
    return v_0 - np.stack([np.zeros_like(z[..., 3]), z[..., 3]], -1)


def generate_x_seq(z, x_0, v_seq, δ):
    x_seq = np.empty((v_seq.shape[0], dim_x))
    x = np.asarray(x_0, dtype=np.float64)
    for t in range(v_seq.shape[0]):
        x = forward_func(z, x, v_seq[t], δ)
        x_seq[t] = x
    return x_seq


def generate_y_seq(z, x_0, v_seq, δ, num_steps_per_obs):
    x_seq = generate_x_seq(z, x_0, v_seq, δ)
    return obs_func(x_seq[num_steps_per_obs - 1:: num_steps_per_obs])


# ---- run-time generator parameters (priors without recompiling) -------------------------------------------------
# The device functor evaluates   z_i = a_i u_i + b_i  (exponentiated where exp_mask_i),   x_0 = v_0 + c + E z
# from 22 run-time values (include/mmd_b200.h: mmd_set_generator_params).  The functions above are the default
# (a = 1, b = 0, exp_mask = [1, 1, 1, 0], c = 0, E[1, 3] = -1); the notebook's priors are another set.
def generator_params(scale=(1.0, 1.0, 1.0, 1.0), shift=(0.0, 0.0, 0.0, 0.0), exp_mask=(1, 1, 1, 0), x0_shift=(0.0, 0.0),
                     x0_z=((0.0, 0.0, 0.0, 0.0), (0.0, 0.0, 0.0, -1.0))):
    """Flat parameter vector for ``BatchedChains(..., generator_params=...)`` / ``mmd_set_generator_params``."""
    p = np.concatenate([np.asarray(scale, float).reshape(4), np.asarray(shift, float).reshape(4),
                        np.asarray(exp_mask, float).reshape(4) != 0, np.asarray(x0_shift, float).reshape(2),
                        np.asarray(x0_z, float).reshape(8)]).astype(np.float64)
    return p


NOTEBOOK_GENERATOR_PARAMS = generator_params(scale=(0.5,) * 4, shift=(-1.0, -2.0, 1.0, 1.0), exp_mask=(1, 1, 0, 0),
                                             x0_shift=(-0.5, -0.5), x0_z=np.zeros((2, 4)))


def make_generators(params):
    """NumPy ``generate_z`` / ``generate_x_0`` for a parameter vector (tagged like the default ones, carrying the
    parameters so that the drop-in system forwards them to the device)."""
    params = np.asarray(params, dtype=np.float64).reshape(22)
    a, b, m, c, E = params[0:4], params[4:8], params[8:12] != 0, params[12:14], params[14:22].reshape(2, 4)

    def generate_z(u):
        lin = a * np.asarray(u)[..., :4] + b
        return np.where(m, np.exp(lin), lin)

    def generate_x_0(z, v_0):
        return np.asarray(v_0) + c + np.asarray(z) @ E.T

    for f in (generate_z, generate_x_0):
        f._mmd_model = "fhn"
        f._mmd_generator_params = params
    return generate_z, generate_x_0
