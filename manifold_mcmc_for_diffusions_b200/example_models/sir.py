"""Susceptible-infected-recovered model with a time-varying (Ornstein-Uhlenbeck log) contact rate: NumPy
mirror of ``sde/example_models/sir.py`` (same names).  ``forward_func`` is the Euler-Maruyama step of the
log-transformed SDE (sir.py:39-70 <- integrators.py:8-14, transforms.py:9-63) in the closed form the CUDA
functor is generated from, including the clip of the first two state components at -500."""
import math

import numpy as np

dim_x = 3
dim_y = 1
dim_w = 3
dim_z = 4
dim_v_0 = 1
dim_v = 3
N = 763.0


def _tag(func):
    func._mmd_model = "sir"
    return func


def _forward_func(z, x, v, δ):
    β, γ, ζ, ϵ = z[0], z[1], z[2], z[3]
    y0, y1, y2 = x[..., 0], x[..., 1], x[..., 2]
    w0, w1, w2 = v[..., 0], v[..., 1], v[..., 2]
    sd = math.sqrt(δ)
    a0 = -(np.exp(y1 + y2) / 2 + np.exp(y0 + y1 + y2)) * np.exp(-y0) / N
    a1 = (-N * β * np.exp(y1) - N * β / 2 - np.exp(y0 + y2) / 2 + np.exp(y0 + y1 + y2)) * np.exp(-y1) / N
    a2 = γ * (ζ - y2)
    b00 = np.exp((-y0 + y1 + y2) / 2) / math.sqrt(N)
    b10 = -np.exp((y0 - y1 + y2) / 2) / math.sqrt(N)
    b11 = np.sqrt(β) * np.exp(-y1 / 2)
    return np.stack([y0 + δ * a0 + sd * b00 * w0, y1 + δ * a1 + sd * (b10 * w0 + b11 * w1),
                     y2 + δ * a2 + sd * ϵ * w2], -1)


@_tag
def forward_func(z, x, v, δ):
    x = np.array(x, dtype=np.float64, copy=True)
    x[..., :2] = np.maximum(x[..., :2], -500.0)
    x_ = _forward_func(z, x, v, δ)
    return np.stack([np.where(x[..., 0] > -500, x_[..., 0], x[..., 0]),
                     np.where(x[..., 1] > -500, x_[..., 1], x[..., 1]), x_[..., 2]], -1)


@_tag
def obs_func(x_seq):
    return np.exp(x_seq[..., 1:2])


@_tag
def generate_z(u):
    return np.stack([np.exp(u[..., 0]), np.exp(u[..., 1]), u[..., 2],
                     np.exp(np.sqrt(0.75) * u[..., 3] + 0.5 * u[..., 1] - 3)], -1)


@_tag
def generate_x_0(z, v_0):
    return np.stack([np.full_like(v_0[..., 0], np.log(762.0)), np.zeros_like(v_0[..., 0]), v_0[..., 0]], -1)


@_tag
def generate_σ_y(u):
    return np.exp(u[..., dim_z])
