"""TEST INFRASTRUCTURE ONLY.  A SymPy-backed stand-in for the few `symnum` names the reference's model definitions use
(`sde/example_models/{fhn,sir}.py`, `sde/integrators.py`, `sde/transforms.py`: `symnum.numpy.array / exp / log / sqrt`,
`symnum.named_array`, `symnum.diffops.symbolic.jacobian / jacobian_vector_product / matrix_hessian_product`,
`symnum.numpify_func`), so that the REFERENCE'S OWN model files can be executed in the build container (symnum 0.1.2
is not installable here) and pin `oracle/models.py` and the generated CUDA functors to them.  Symbolic arrays are
NumPy object arrays of SymPy expressions; `numpify_func` lambdifies onto torch (float64).  Nothing under the product
package imports this module."""
import math
import sys
import types

import numpy as onp
import sympy
import torch


class SymbolicArray(onp.ndarray):
    """NumPy object array of SymPy expressions with the methods the reference calls on symnum arrays."""

    def __new__(cls, obj):
        arr = onp.array(obj, dtype=object)
        flat = [sympy.sympify(e) for e in arr.reshape(-1)]
        out = onp.empty(len(flat), dtype=object)
        out[:] = flat
        return out.reshape(arr.shape).view(cls)

    def diff(self, *variables):
        out = self
        for var in variables:
            if isinstance(var, onp.ndarray):
                out = SymbolicArray([[sympy.diff(e, v) for v in var.reshape(-1)] for e in out.reshape(-1)]).reshape(
                    out.shape + var.shape)
            else:
                out = SymbolicArray([sympy.diff(e, var) for e in out.reshape(-1)]).reshape(out.shape)
        return out

    def subs(self, *args):
        return SymbolicArray([e.subs(*args) for e in self.reshape(-1)]).reshape(self.shape)

    def simplify(self):
        return SymbolicArray([sympy.simplify(e) for e in self.reshape(-1)]).reshape(self.shape)


def _elementwise(fn):
    def f(x):
        if isinstance(x, onp.ndarray):
            return SymbolicArray([fn(e) for e in x.reshape(-1)]).reshape(x.shape)
        return fn(sympy.sympify(x))

    return f


def named_array(name, shape):
    if shape is None or shape == ():
        return sympy.Symbol(name)
    if isinstance(shape, int):
        shape = (shape,)
    syms = [sympy.Symbol(f"{name}[{', '.join(map(str, idx))}]") for idx in onp.ndindex(*shape)]
    return SymbolicArray(syms).reshape(shape)


def _jacobian(func, wrt=0):
    def jac(*args):
        f, x = onp.asarray(func(*args), dtype=object), args[wrt]
        J = [[sympy.diff(fi, xj) for xj in x.reshape(-1)] for fi in f.reshape(-1)]
        return SymbolicArray(J).reshape(f.shape + x.shape)

    return jac


def _jacobian_vector_product(func, wrt=0):
    def jvp(*args):
        J = _jacobian(func, wrt)(*args)
        return lambda v: SymbolicArray(J @ onp.asarray(v, dtype=object))

    return jvp


def _matrix_hessian_product(func, wrt=0):
    def mhp(*args):
        f, x = onp.asarray(func(*args), dtype=object), args[wrt]
        xs = list(x.reshape(-1))

        def apply(M):
            M = onp.asarray(M, dtype=object)
            return SymbolicArray([sum(sympy.diff(fi, xs[k], xs[l]) * M[k, l] for k in range(len(xs)) for l in range(len(xs)))
                                  for fi in f.reshape(-1)]).reshape(f.shape)

        return apply

    return mhp


def _num(fn_t, fn_m):
    return lambda a: fn_t(a) if isinstance(a, torch.Tensor) else fn_m(a)


_LAMBDIFY_FUNCS = {"exp": _num(torch.exp, math.exp), "log": _num(torch.log, math.log),
                   "sqrt": _num(torch.sqrt, math.sqrt), "Abs": _num(torch.abs, abs)}


def numpify_func(func, *arg_shapes, numpy_module=None, **_):
    """symnum.numpify_func: a numeric function of arrays with the given shapes (None: scalar) from a function of symbolic
    arrays; here onto torch float64 (the `numpy_module` of the reference, jax.numpy, is the torch stand-in)."""
    sym_args = [named_array(f"arg{i}", s) for i, s in enumerate(arg_shapes)]
    out = onp.asarray(func(*sym_args), dtype=object)
    flat_syms = []
    for a in sym_args:
        flat_syms.extend(list(a.reshape(-1)) if isinstance(a, onp.ndarray) else [a])
    f = sympy.lambdify(flat_syms, [sympy.sympify(e) for e in out.reshape(-1)], modules=[_LAMBDIFY_FUNCS, "math"])

    def numeric(*vals):
        flat = []
        for v, s in zip(vals, arg_shapes):
            if s is None or s == ():
                flat.append(v)
            else:
                v = v if isinstance(v, torch.Tensor) else torch.as_tensor(onp.asarray(v), dtype=torch.float64)
                flat.extend(v.reshape(-1)[i] for i in range(int(onp.prod(s))))
        res = f(*flat)
        return torch.stack([r if isinstance(r, torch.Tensor) else torch.as_tensor(float(r), dtype=torch.float64)
                            for r in res]).reshape(out.shape)

    return numeric


def install():
    """Register symnum, symnum.numpy, symnum.diffops, symnum.diffops.symbolic in sys.modules (idempotent)."""
    if "symnum" in sys.modules and not getattr(sys.modules["symnum"], "_mmd_sympy_shim", False):
        raise RuntimeError("a real symnum is imported: the SymPy stand-in is only for containers without it")
    symnum = types.ModuleType("symnum")
    symnum._mmd_sympy_shim = True
    snp = types.ModuleType("symnum.numpy")
    snp.array = SymbolicArray
    snp.exp, snp.log, snp.sqrt = _elementwise(sympy.exp), _elementwise(sympy.log), _elementwise(sympy.sqrt)
    snp.SymbolicArray = SymbolicArray
    diffops = types.ModuleType("symnum.diffops")
    symbolic = types.ModuleType("symnum.diffops.symbolic")
    symbolic.jacobian, symbolic.jacobian_vector_product = _jacobian, _jacobian_vector_product
    symbolic.matrix_hessian_product = _matrix_hessian_product
    diffops.symbolic = symbolic
    symnum.numpy, symnum.diffops = snp, diffops
    symnum.named_array, symnum.numpify_func = named_array, numpify_func
    for name, mod in (("symnum", symnum), ("symnum.numpy", snp), ("symnum.diffops", diffops),
                      ("symnum.diffops.symbolic", symbolic)):
        sys.modules[name] = mod
    return symnum
