"""TEST INFRASTRUCTURE ONLY.  A torch-backed stand-in for the few `jax` names the reference's
`sde/mici_extensions.py` uses, so that the REFERENCE'S OWN SOURCE FILE can be executed in the build container
(where jax 0.2.21 / jaxlib 0.1.71 cannot be installed) and pin the restated oracle and the golden vectors at trace
level: `jax.numpy` -> float64 torch ops, `jax.jit` -> identity (NumPy arguments become tensors at the boundary),
`jax.vmap / grad / value_and_grad / jacrev / jacobian` -> `torch.func`, `lax.scan / while_loop` -> Python loops,
`jax.scipy.linalg.cho_solve / lu_factor / lu_solve` -> `torch.linalg`.  Only what the reference file touches is
provided (reference lines cited at each piece); nothing under the product package imports this module.

    from oracle.jax_torch_shim import install
    install()            # registers jax, jax.numpy, jax.numpy.linalg, jax.scipy.linalg, jax.lax, ... in sys.modules
"""
import functools
import sys
import types

import numpy as onp
import torch

F64 = torch.float64


def _t(x):
    """NumPy array / scalar / (nested) sequence -> float64 (or integer / bool) tensor; tensors pass through."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)):
        if len(x) and any(isinstance(e, (torch.Tensor, list, tuple)) for e in x):
            return torch.stack([_t(e) for e in x])
        x = onp.asarray(x)
    if isinstance(x, onp.ndarray):
        t = torch.from_numpy(onp.ascontiguousarray(x))
        return t.to(F64) if t.dtype in (torch.float32, torch.float16) else t
    if isinstance(x, (bool, onp.bool_)):
        return torch.tensor(bool(x))
    if isinstance(x, (int, onp.integer)):
        return torch.tensor(int(x))
    if isinstance(x, (float, onp.floating)):
        return torch.tensor(float(x), dtype=F64)
    return x


def _tree(x):
    """Convert the NumPy leaves of a (nested tuple / list) argument to tensors (the jit boundary)."""
    if isinstance(x, onp.ndarray):
        return _t(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_tree(e) for e in x)
    return x


class _At:
    """`x.at[idx].set(v)` of jax arrays (mici_extensions.py: two uses) for tensors."""

    def __init__(self, x):
        self.x, self.idx = x, None

    def __getitem__(self, idx):
        self.idx = idx
        return self

    def set(self, v):
        y = self.x.clone()
        y[self.idx] = _t(v)
        return y

    def add(self, v):
        y = self.x.clone()
        y[self.idx] = y[self.idx] + _t(v)
        return y


def _axis_kw(axis):
    return {} if axis is None else {"dim": axis}


def _make_numpy():
    np = types.ModuleType("jax.numpy")
    np.inf, np.pi, np.newaxis = float("inf"), onp.pi, None
    np.float64, np.int32, np.int64, np.bool_ = F64, torch.int32, torch.int64, torch.bool
    np.DeviceArray = np.ndarray = torch.Tensor
    np.asarray = np.array = lambda x, dtype=None: _t(x) if dtype is None else _t(x).to(dtype)
    np.zeros = lambda shape, dtype=None: torch.zeros(shape, dtype=dtype or F64)
    np.ones = lambda shape, dtype=None: torch.ones(shape, dtype=dtype or F64)
    np.zeros_like = lambda x, dtype=None: torch.zeros_like(_t(x), dtype=dtype)
    np.ones_like = lambda x, dtype=None: torch.ones_like(_t(x), dtype=dtype)
    np.identity = np.eye = lambda n, dtype=None: torch.eye(n, dtype=dtype or F64)
    np.arange = lambda *a, **k: torch.arange(*a, **k)
    np.concatenate = lambda seq, axis=0: torch.cat([_t(e) for e in seq], dim=axis)
    np.vstack = lambda seq: torch.vstack([_t(e) for e in seq])
    np.hstack = lambda seq: torch.hstack([_t(e) for e in seq])
    np.stack = lambda seq, axis=0: torch.stack([_t(e) for e in seq], dim=axis)
    np.reshape = lambda x, shape: _t(x).reshape(shape)
    np.einsum = lambda subs, *ops: torch.einsum(subs, *[_t(o) for o in ops])
    np.outer = lambda a, b: torch.outer(_t(a), _t(b))
    np.dot = lambda a, b: _t(a) @ _t(b)
    np.matmul = lambda a, b: _t(a) @ _t(b)
    np.where = lambda c, a, b: torch.where(_t(c), _t(a), _t(b))
    np.diag_indices = lambda n: (torch.arange(n), torch.arange(n))
    np.diag = lambda x: torch.diag(_t(x))
    np.clip = lambda x, a_min=None, a_max=None: torch.clamp(_t(x), min=a_min, max=a_max)   # sir.py:63
    np.sum = lambda x, axis=None: torch.sum(_t(x), **_axis_kw(axis))
    np.mean = lambda x, axis=None: torch.mean(_t(x), **_axis_kw(axis))
    np.max = lambda x, axis=None: torch.max(_t(x)) if axis is None else torch.max(_t(x), dim=axis).values
    np.min = lambda x, axis=None: torch.min(_t(x)) if axis is None else torch.min(_t(x), dim=axis).values
    np.all = lambda x: torch.all(_t(x))
    np.any = lambda x: torch.any(_t(x))
    for name in ("abs", "log", "exp", "sqrt", "sin", "cos", "tanh", "isnan", "isfinite", "logical_not", "square"):
        setattr(np, name, (lambda f: lambda x: f(_t(x)))(getattr(torch, name)))
    for name in ("logical_or", "logical_and", "maximum", "minimum"):
        setattr(np, name, (lambda f: lambda a, b: f(_t(a), _t(b)))(getattr(torch, name)))
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.cholesky = lambda a: torch.linalg.cholesky(_t(a))
    linalg.solve = lambda a, b: torch.linalg.solve(_t(a), _t(b))
    linalg.slogdet = lambda a: tuple(torch.linalg.slogdet(_t(a)))
    linalg.norm = lambda x, ord=None: torch.linalg.norm(_t(x), ord=ord)
    # np.linalg.lstsq(A, b)[0] on a square full-rank system (mici_extensions.py:1512-1516) == solve
    linalg.lstsq = lambda a, b, rcond=None: (torch.linalg.solve(_t(a), _t(b)), None, None, None)
    np.linalg = linalg
    return np, linalg


def _make_scipy_linalg():
    sla = types.ModuleType("jax.scipy.linalg")

    def cho_solve(c_and_lower, b):      # :915-942 (lmult_by_inv_gram and friends)
        c, lower = c_and_lower
        return torch.cholesky_solve(_t(b).unsqueeze(-1) if _t(b).dim() == _t(c).dim() - 1 else _t(b), _t(c),
                                    upper=not lower).reshape(_t(b).shape)

    def cholesky(a, lower=False):
        L = torch.linalg.cholesky(_t(a))
        return L if lower else L.mT

    def lu_factor(a):                   # :745-763
        lu, piv = torch.linalg.lu_factor(_t(a))
        return lu, piv

    def lu_solve(lu_and_piv, b, trans=0):  # :944-981
        lu, piv = lu_and_piv
        b = _t(b)
        vec = b.dim() == lu.dim() - 1
        bb = b.unsqueeze(-1) if vec else b
        if trans == 0:
            x = torch.linalg.lu_solve(lu, piv, bb)
        else:
            x = torch.linalg.lu_solve(lu, piv, bb, adjoint=True)
        return x.squeeze(-1) if vec else x

    def solve_triangular(a, b, lower=False, trans=0):
        a, b = _t(a), _t(b)
        vec = b.dim() == a.dim() - 1
        bb = b.unsqueeze(-1) if vec else b
        if trans in (1, "T"):
            a, lower = a.mT, not lower
        x = torch.linalg.solve_triangular(a, bb, upper=not lower)
        return x.squeeze(-1) if vec else x

    sla.cho_solve, sla.cholesky, sla.lu_factor, sla.lu_solve, sla.solve_triangular = (
        cho_solve, cholesky, lu_factor, lu_solve, solve_triangular)
    return sla


def _make_lax():
    lax = types.ModuleType("jax.lax")

    def _index(xs, i):
        if xs is None:
            return None
        if isinstance(xs, (tuple, list)):
            return type(xs)(_index(e, i) for e in xs)
        return _t(xs)[i]

    def _length(xs):
        if isinstance(xs, (tuple, list)):
            return _length(xs[0])
        return _t(xs).shape[0]

    def _stack(ys):
        if ys[0] is None:
            return None
        if isinstance(ys[0], (tuple, list)):
            return type(ys[0])(_stack([y[j] for y in ys]) for j in range(len(ys[0])))
        return torch.stack(ys)

    def scan(f, init, xs, length=None):   # :178, :396-402, :1611, :1725
        n = length if xs is None else _length(xs)
        carry, ys = init, []
        for i in range(n):
            carry, y = f(carry, _index(xs, i))
            ys.append(y)
        return carry, _stack(ys)

    def while_loop(cond_fun, body_fun, init_val):   # :1057, :1129 (projection solver loops)
        val = init_val
        while bool(cond_fun(val)):
            val = body_fun(val)
        return val

    def map_(f, xs):
        return _stack([f(_index(xs, i)) for i in range(_length(xs))])

    lax.scan, lax.while_loop, lax.map = scan, while_loop, map_
    lax.cond = lambda pred, tf, ff, *ops: tf(*ops) if bool(pred) else ff(*ops)
    lax.select = lambda pred, a, b: torch.where(_t(pred), _t(a), _t(b))   # sir.py:66-70
    return lax


def _argnums(a):
    return tuple(a) if isinstance(a, (tuple, list)) else a


# torch.func's has_aux wants a tensors-only pytree; the reference's auxiliary outputs contain None entries (absent
# blocks): they travel through the transform as empty tensors
def _hide_none(x):
    if x is None:
        return torch.zeros(0, dtype=torch.int8)
    if isinstance(x, (tuple, list)):
        return type(x)(_hide_none(e) for e in x)
    return x


def _show_none(x):
    if isinstance(x, torch.Tensor) and x.dtype == torch.int8 and x.numel() == 0:
        return None
    if isinstance(x, (tuple, list)):
        return type(x)(_show_none(e) for e in x)
    return x


def install(force=False):
    """Register the stand-in modules (idempotent).  Refuses to shadow a real jax unless `force`."""
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "_mmd_torch_shim", False) and not force:
        raise RuntimeError("a real jax is imported: the torch stand-in is only for containers without it")
    torch.set_default_dtype(F64)
    if not hasattr(torch.Tensor, "at"):
        torch.Tensor.at = property(lambda self: _At(self))
    if not hasattr(onp, "product"):      # removed in NumPy 2; the reference calls it on shape tuples (:49)
        onp.product = onp.prod
    jax = types.ModuleType("jax")
    jax._mmd_torch_shim = True
    np, nla = _make_numpy()
    sla, lax = _make_scipy_linalg(), _make_lax()

    def jit(f=None, static_argnums=None, static_argnames=None, **_):
        if f is None:
            return functools.partial(jit, static_argnums=static_argnums, static_argnames=static_argnames)

        @functools.wraps(f)
        def wrapped(*args, **kwargs):
            return f(*[_tree(a) for a in args], **{k: _tree(v) for k, v in kwargs.items()})

        return wrapped

    def vmap(f, in_axes=0, out_axes=0):
        in_dims = tuple(in_axes) if isinstance(in_axes, (list, tuple)) else in_axes
        g = torch.func.vmap(f, in_dims=in_dims, out_dims=out_axes)
        return lambda *args: g(*[_tree(a) for a in args])

    def _aux_safe(f, has_aux):
        if not has_aux:
            return f

        def g(*args, **kwargs):
            val, aux = f(*args, **kwargs)
            return val, _hide_none(aux)

        return g

    def grad(f, argnums=0, has_aux=False):
        g = torch.func.grad(_aux_safe(f, has_aux), argnums=_argnums(argnums), has_aux=has_aux)
        if not has_aux:
            return lambda *args, **kw: g(*[_tree(a) for a in args], **kw)

        def wrapped(*args, **kwargs):
            gr, aux = g(*[_tree(a) for a in args], **kwargs)
            return gr, _show_none(aux)

        return wrapped

    def value_and_grad(f, argnums=0, has_aux=False):
        g = torch.func.grad_and_value(_aux_safe(f, has_aux), argnums=_argnums(argnums), has_aux=has_aux)

        def wrapped(*args, **kwargs):
            gr, val = g(*[_tree(a) for a in args], **kwargs)   # jax order: (value [, aux]), grad
            if has_aux:
                val = (val[0], _show_none(val[1]))
            return val, gr

        return wrapped

    def jacrev(f, argnums=0):
        return torch.func.jacrev(f, argnums=_argnums(argnums))

    def jacfwd(f, argnums=0):
        return torch.func.jacfwd(f, argnums=_argnums(argnums))

    jax.jit, jax.vmap, jax.grad, jax.value_and_grad = jit, vmap, grad, value_and_grad
    jax.jacrev, jax.jacobian, jax.jacfwd = jacrev, jacrev, jacfwd
    jax.numpy, jax.lax = np, lax
    config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.config = config
    scipy = types.ModuleType("jax.scipy")
    scipy.linalg = sla
    jax.scipy = scipy
    experimental = types.ModuleType("jax.experimental")
    optimizers = types.ModuleType("jax.experimental.optimizers")

    def adam(step_size, b1=0.9, b2=0.999, eps=1e-8):   # :1679-1801 (the Adam initialiser); standard update rule
        def init(x0):
            return (_t(x0), torch.zeros_like(_t(x0)), torch.zeros_like(_t(x0)))

        def update(i, g, state):
            x, m, v = state
            g = _t(g)
            m = (1 - b1) * g + b1 * m
            v = (1 - b2) * g * g + b2 * v
            mhat, vhat = m / (1 - b1 ** (i + 1)), v / (1 - b2 ** (i + 1))
            return (x - step_size * mhat / (torch.sqrt(vhat) + eps), m, v)

        return init, update, (lambda state: state[0])

    optimizers.adam = adam
    experimental.optimizers = optimizers
    jax.experimental = experimental
    api = types.ModuleType("jax.api")
    api.jit, api.vmap, api.grad, api.value_and_grad, api.jacrev, api.jacobian = jit, vmap, grad, value_and_grad, jacrev, jacrev
    jax.api = api
    for name, mod in (("jax", jax), ("jax.numpy", np), ("jax.numpy.linalg", nla), ("jax.scipy", scipy),
                      ("jax.scipy.linalg", sla), ("jax.lax", lax), ("jax.config", config),
                      ("jax.experimental", experimental), ("jax.experimental.optimizers", optimizers),
                      ("jax.api", api)):
        sys.modules[name] = mod
    return jax
