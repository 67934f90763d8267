"""FitzHugh-Nagumo model with the prior parametrisation of the reference's notebook
(``FitzHugh-Nagumo_example.ipynb`` cells 7-20): the strong-order-1.5 step of :mod:`.fhn`, with
``generate_z(u) = [exp(.5 u0 - 1), exp(.5 u1 - 2), .5 u2 + 1, .5 u3 + 1]`` and
``generate_x_0(z, v_0) = [-.5, -.5] + v_0``."""
import numpy as np

from . import fhn as _fhn

dim_x, dim_w, dim_z, dim_v_0, dim_v = 2, 1, 4, 2, 2


def _tag(func):
    func._mmd_model = "fhn_notebook"
    return func


@_tag
def forward_func(z, x, v, δ):
    return _fhn.forward_func.__wrapped__(z, x, v, δ) if hasattr(_fhn.forward_func, "__wrapped__") else _step(z, x, v, δ)


def _step(z, x, v, δ):
    import math

    σ, ε, γ, β = z[0], z[1], z[2], z[3]
    x0, x1 = x[..., 0], x[..., 1]
    v0, v1 = v[..., 0], v[..., 1]
    P = x0 ** 3 - x0 + x1
    Q = β + γ * x0 - x1
    a0 = -P / ε
    dζ = δ ** 1.5 * (v0 + v1 / math.sqrt(3.0)) / 2
    f0 = x0 + δ * a0 + (δ ** 2 / 2) * (((1 - 3 * x0 ** 2) / ε) * a0 - Q / ε) - (σ / ε) * dζ
    f1 = x1 + δ * Q + σ * math.sqrt(δ) * v0 + (δ ** 2 / 2) * (γ * a0 - Q) - σ * dζ
    return np.stack([f0, f1], -1)


@_tag
def obs_func(x_seq):
    return x_seq[..., 0:1]


@_tag
def generate_z(u):
    return np.stack([np.exp(0.5 * u[..., 0] - 1), np.exp(0.5 * u[..., 1] - 2), 0.5 * u[..., 2] + 1,
                     0.5 * u[..., 3] + 1], -1)


@_tag
def generate_x_0(z, v_0):
    return np.array([-0.5, -0.5]) + v_0


def generate_from_model(q, δ, num_steps_per_obs):
    """Notebook cell 20: parameters, state sequence and observations generated from the latent vector q."""
    u, v_0, v_r = q[:dim_z], q[dim_z: dim_z + dim_x], q[dim_z + dim_x:]
    z = generate_z(u)
    x = generate_x_0(z, v_0)
    v_seq = v_r.reshape(-1, dim_v)
    x_seq = np.empty((v_seq.shape[0], dim_x))
    for t in range(v_seq.shape[0]):
        x = _step(z, x, v_seq[t], δ)
        x_seq[t] = x
    return x_seq, obs_func(x_seq[num_steps_per_obs - 1:: num_steps_per_obs]), z, generate_x_0(z, v_0)
