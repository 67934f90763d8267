"""Example diffusion models for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Restates ``sde/example_models/fhn.py`` and ``sde/example_models/sir.py`` of the reference with
``torch`` (float64) in place of ``jax.numpy``.  The reference builds its one-step maps by running
``sde/integrators.py`` through SymNum (not installed here); ``derive_*_step`` below re-applies the
same operator definitions with plain SymPy and the hard-coded expressions used at run time are
checked against that derivation in ``tests/test_oracle_models.py``.

Nothing in the product package may import this module.
"""

import math
from types import SimpleNamespace

import torch

# ---------------------------------------------------------------------------------------------
# SymPy re-derivation of the reference step maps (sde/integrators.py:8-14, 46-63, 95-149;
# sde/transforms.py:9-63).  Only used by tests / code generation, never at oracle run time.
# ---------------------------------------------------------------------------------------------


def _jvp(func, x, z, vec):
    import sympy as sp

    return func(x, z).jacobian(x) * vec


def _mhp(func, x, z, M):
    import sympy as sp

    f = func(x, z)
    n = len(x)
    return sp.Matrix(
        [
            sum(sp.hessian(f[i], x)[j, k] * M[j, k] for j in range(n) for k in range(n))
            for i in range(f.shape[0])
        ]
    )


def derive_fhn_step(simplify=True):
    """Strong-order-1.5 (additive noise) step for FHN: integrators.py:46-63 on fhn.py:17-24."""
    import sympy as sp

    x0, x1, v0, v1 = sp.symbols("x0 x1 v0 v1", real=True)
    s, e, g, b = sp.symbols("sigma epsilon gamma beta", real=True)
    d = sp.symbols("delta", positive=True)
    x = sp.Matrix([x0, x1])
    v = sp.Matrix([v0, v1])
    z = (s, e, g, b)

    def drift(x, z):
        s, e, g, b = z
        return sp.Matrix([(x[0] - x[0] ** 3 - x[1]) / e, g * x[0] - x[1] + b])

    def diff(x, z):
        s, e, g, b = z
        return sp.Matrix([[0], [s]])

    a = drift(x, z)
    B = diff(x, z)
    # diffusion_operator(drift, diff)(drift): integrators.py:95-123
    L0a = _jvp(drift, x, z, a) + _mhp(drift, x, z, B * B.T) / 2
    dim_noise = 1
    dw = sp.sqrt(d) * v[:dim_noise, :]
    dzeta = d * sp.sqrt(d) * (v[:dim_noise, :] + v[dim_noise:, :] / sp.sqrt(3)) / 2
    xn = x + d * a + B * dw + (d ** 2 / 2) * L0a
    for j in range(dim_noise):
        # Lj_operator(diff, j)(drift): integrators.py:126-149
        xn = xn + _jvp(drift, x, z, B[:, j]) * dzeta[j]
    if simplify:
        xn = sp.Matrix([sp.simplify(xn[i]) for i in range(2)])
    syms = dict(x=(x0, x1), v=(v0, v1), z=z, delta=d)
    return xn, syms


def derive_sir_step(simplify=True):
    """Euler-Maruyama on the log-transformed SIR SDE: sir.py:19-51, transforms.py:9-63."""
    import sympy as sp

    y0, y1, y2, w0, w1, w2 = sp.symbols("y0 y1 y2 w0 w1 w2", real=True)
    be, ga, ze, ep = sp.symbols("beta gamma zeta epsilon", real=True)
    d = sp.symbols("delta", positive=True)
    N = 763
    X = sp.symbols("X0 X1 X2", positive=True)
    xs = sp.Matrix(X)
    z = (be, ga, ze, ep)

    def drift(x, z):
        al = sp.exp(x[2])
        b_, g_, z_, e_ = z
        return sp.Matrix(
            [-al * x[0] * x[1] / N, al * x[0] * x[1] / N - b_ * x[1], g_ * (z_ - x[2])]
        )

    def diffc(x, z):
        al = sp.exp(x[2])
        b_, g_, z_, e_ = z
        r = sp.sqrt(al * x[0] * x[1] / N)
        return sp.Matrix([[r, 0, 0], [-r, sp.sqrt(b_ * x[1]), 0], [0, 0, e_]])

    fwd = sp.Matrix([sp.log(X[0]), sp.log(X[1]), X[2]])
    a = drift(xs, z)
    B = diffc(xs, z)
    J = fwd.jacobian(xs)
    BBt = B * B.T
    hess_term = sp.Matrix(
        [
            sum(sp.hessian(fwd[i], xs)[j, k] * BBt[j, k] for j in range(3) for k in range(3))
            for i in range(3)
        ]
    )
    y = sp.Matrix([y0, y1, y2])
    back = {X[0]: sp.exp(y0), X[1]: sp.exp(y1), X[2]: y2}
    a_y = (J * a + hess_term / 2).subs(back)
    B_y = (J * B).subs(back)
    if simplify:
        a_y = a_y.applyfunc(sp.simplify)
        B_y = B_y.applyfunc(sp.simplify)
    w = sp.Matrix([w0, w1, w2])
    yn = y + d * a_y + sp.sqrt(d) * B_y * w
    syms = dict(x=(y0, y1, y2), v=(w0, w1, w2), z=z, delta=d)
    return yn, syms


# ---------------------------------------------------------------------------------------------
# FitzHugh-Nagumo (sde/example_models/fhn.py)
# ---------------------------------------------------------------------------------------------

_SQRT3 = math.sqrt(3.0)


def fhn_forward_func(z, x, v, δ):
    """One strong-order-1.5 step, fhn.py:27-34 (expression = SymPy 1.14 `simplify` output of
    `derive_fhn_step`, evaluated in the order printed)."""
    σ, ε, γ, β = z[0], z[1], z[2], z[3]
    x0, x1 = x[0], x[1]
    v0, v1 = v[0], v[1]
    P = x0 ** 3 - x0 + x1
    Q = β + γ * x0 - x1
    noise = δ ** 1.5 * σ * (3 * v0 + _SQRT3 * v1)
    f0 = (
        -3 * δ ** 2 * (ε * Q - (3 * x0 ** 2 - 1) * P)
        + 6 * ε ** 2 * x0
        - ε * (noise + 6 * δ * P)
    ) / (6 * ε ** 2)
    f1 = (
        -3 * δ ** 2 * (ε * Q + γ * P)
        + ε * (-noise + 6 * math.sqrt(δ) * σ * v0 + 6 * δ * Q + 6 * x1)
    ) / (6 * ε)
    return torch.stack([f0, f1])


def fhn_obs_func(x_seq):  # fhn.py:37-38
    return x_seq[..., 0:1]


def fhn_generate_z(u):  # fhn.py:41-43  [σ, ϵ, γ, β]
    return torch.stack([torch.exp(u[0]), torch.exp(u[1]), torch.exp(u[2]), u[3]])


def fhn_generate_σ_y(u):  # fhn.py:46-47
    return torch.exp(u[4])


def fhn_generate_x_0(z, v_0):  # fhn.py:50-51
    return v_0 - torch.stack([torch.zeros_like(z[3]), z[3]])


def fhn_generate_x_seq(z, x_0, v_seq, δ):  # fhn.py:54-60
    xs = []
    x = x_0
    for t in range(v_seq.shape[0]):
        x = fhn_forward_func(z, x, v_seq[t], δ)
        xs.append(x)
    return torch.stack(xs)


def fhn_generate_y_seq(z, x_0, v_seq, δ, num_steps_per_obs):  # fhn.py:63-65
    x_seq = fhn_generate_x_seq(z, x_0, v_seq, δ)
    return fhn_obs_func(x_seq[num_steps_per_obs - 1 :: num_steps_per_obs])


fhn = SimpleNamespace(
    name="fhn",
    dim_x=2,
    dim_w=1,
    dim_z=4,
    dim_v_0=2,
    dim_v=2,
    forward_func=fhn_forward_func,
    obs_func=fhn_obs_func,
    generate_z=fhn_generate_z,
    generate_σ_y=fhn_generate_σ_y,
    generate_x_0=fhn_generate_x_0,
    generate_x_seq=fhn_generate_x_seq,
    generate_y_seq=fhn_generate_y_seq,
)


def fhn_simulate_y_seq_numpy(z, x_0, v_seq, δ, num_steps_per_obs):
    """NumPy scalar-loop version of `fhn.generate_y_seq` for the 10,000 steps/obs data simulation
    (fhn_model_noiseless_obs_chmc_experiment.py:84-93); same arithmetic as `fhn_forward_func`."""
    import numpy as onp

    σ, ε, γ, β = (float(t) for t in z)
    x0, x1 = float(x_0[0]), float(x_0[1])
    sd = math.sqrt(δ)
    d15 = δ ** 1.5
    n = v_seq.shape[0]
    ys = onp.empty((n // num_steps_per_obs, 1))
    for t in range(n):
        v0, v1 = v_seq[t, 0], v_seq[t, 1]
        P = x0 ** 3 - x0 + x1
        Q = β + γ * x0 - x1
        noise = d15 * σ * (3 * v0 + _SQRT3 * v1)
        f0 = (
            -3 * δ ** 2 * (ε * Q - (3 * x0 ** 2 - 1) * P) + 6 * ε ** 2 * x0 - ε * (noise + 6 * δ * P)
        ) / (6 * ε ** 2)
        f1 = (
            -3 * δ ** 2 * (ε * Q + γ * P) + ε * (-noise + 6 * sd * σ * v0 + 6 * δ * Q + 6 * x1)
        ) / (6 * ε)
        x0, x1 = f0, f1
        if (t + 1) % num_steps_per_obs == 0:
            ys[(t + 1) // num_steps_per_obs - 1, 0] = x0
    return ys


# ---------------------------------------------------------------------------------------------
# SIR with OU log-contact-rate (sde/example_models/sir.py)
# ---------------------------------------------------------------------------------------------

SIR_N = 763.0


def _sir_forward_func_raw(z, x, v, δ):
    """Euler-Maruyama step on [log S, log I, log-rate], sir.py:39-51 (closed form of
    `derive_sir_step`)."""
    β, γ, ζ, ϵ = z[0], z[1], z[2], z[3]
    y0, y1, y2 = x[0], x[1], x[2]
    w0, w1, w2 = v[0], v[1], v[2]
    N = SIR_N
    sd = math.sqrt(δ)
    a0 = -(torch.exp(y1 + y2) / 2 + torch.exp(y0 + y1 + y2)) * torch.exp(-y0) / N
    a1 = (
        (-N * β * torch.exp(y1) - N * β / 2 - torch.exp(y0 + y2) / 2 + torch.exp(y0 + y1 + y2))
        * torch.exp(-y1)
        / N
    )
    a2 = γ * (ζ - y2)
    b00 = torch.exp((-y0 + y1 + y2) / 2) / math.sqrt(N)
    b10 = -torch.exp((y0 - y1 + y2) / 2) / math.sqrt(N)
    b11 = torch.sqrt(β) * torch.exp(-y1 / 2)
    f0 = y0 + δ * a0 + sd * b00 * w0
    f1 = y1 + δ * a1 + sd * (b10 * w0 + b11 * w1)
    f2 = y2 + δ * a2 + sd * ϵ * w2
    return torch.stack([f0, f1, f2])


def sir_forward_func(z, x, v, δ):  # sir.py:54-70 (clip before, select after)
    x01 = torch.clamp(x[:2], min=-500.0)
    xc = torch.cat([x01, x[2:]])
    x_ = _sir_forward_func_raw(z, xc, v, δ)
    return torch.stack(
        [
            torch.where(xc[0] > -500, x_[0], xc[0]),
            torch.where(xc[1] > -500, x_[1], xc[1]),
            x_[2],
        ]
    )


def sir_obs_func(x_seq):  # sir.py:73-74
    return torch.exp(x_seq[..., 1:2])


def sir_generate_z(u):  # sir.py:77-85
    return torch.stack(
        [
            torch.exp(u[0]),
            torch.exp(u[1]),
            u[2],
            torch.exp(math.sqrt(0.75) * u[3] + 0.5 * u[1] - 3),
        ]
    )


def sir_generate_x_0(z, v_0):  # sir.py:88-89
    c = torch.tensor([math.log(762.0), math.log(1.0)], dtype=v_0.dtype)
    return torch.cat([c, v_0[0:1]])


def sir_generate_σ_y(u):  # sir.py:92-93
    return torch.exp(u[4])


sir = SimpleNamespace(
    name="sir",
    dim_x=3,
    dim_y=1,
    dim_w=3,
    dim_z=4,
    dim_v_0=1,
    dim_v=3,
    forward_func=sir_forward_func,
    obs_func=sir_obs_func,
    generate_z=sir_generate_z,
    generate_x_0=sir_generate_x_0,
    generate_σ_y=sir_generate_σ_y,
)
