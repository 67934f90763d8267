"""Import stand-ins that let the reference's scripts run unchanged on top of this package when their Python-only
dependencies are absent (none of this is on the GPU hot path; each stand-in is installed only if the real package
cannot be imported):

  jax          scripts/utils.py:12 (`jax.config.update`), fhn_model_noiseless_obs_chmc_operation_times.py:13,18,33,35
               (`jit`, `lax.map`, `jax.numpy.DeviceArray`).  `jit(f)` is `f`; `lax.map(f, xs)` calls `f(xs)` ONCE with
               the whole batch -- the private handles of ConditionedDiffusionConstrainedSystem take a leading batch
               axis and run one device launch for all states; `DeviceArray` is the handles' result type.
  arviz        scripts/utils.py:369 (`arviz.summary(traces, var_names=...)`): rank-normalised split-R-hat / bulk ESS
               from `diagnostics.py`, returned as a pandas DataFrame with ArviZ's column names.
  matplotlib   scripts/utils.py:11 (imported at module level, only used by the plotting helpers): an object that
               raises on use.

    from manifold_mcmc_for_diffusions_b200 import install_reference_aliases
    from manifold_mcmc_for_diffusions_b200.compat_shims import install_import_shims
    install_reference_aliases(); install_import_shims()
    runpy.run_path("scripts/fhn_model_noiseless_obs_chmc_experiment.py", run_name="__main__")
"""

import importlib
import sys
import types

import numpy as np


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except ImportError:
        return True


def _make_jax():
    from .mici_extensions import DeviceArray

    jax = types.ModuleType("jax")
    config = types.ModuleType("jax.config")
    config.values = {}
    config.update = lambda key, value: config.values.__setitem__(key, value)
    # `import jax.config` binds the submodule, `jax.config.update(...)` is then called on it (utils.py:19-23);
    # newer JAX exposes the same `update` on an object: both spellings land here
    config.config = config
    jax.config = config

    def jit(func=None, **kwargs):
        if func is None:
            return lambda f: f
        return func

    lax = types.ModuleType("jax.lax")

    def lax_map(func, xs):
        """The mapped function receives the stacked inputs in one call (the device handles are batched)."""
        return func(xs)

    lax.map = lax_map
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            setattr(jnp, name, getattr(np, name))
    jnp.DeviceArray = DeviceArray
    jax.jit, jax.lax, jax.numpy = jit, lax, jnp
    return {"jax": jax, "jax.config": config, "jax.lax": lax, "jax.numpy": jnp}


def _make_arviz():
    arviz = types.ModuleType("arviz")

    def summary(data, var_names=None, **kwargs):
        """traces: {var: list over chains of arrays [n_iter] or [n_iter, k]} as Mici returns them."""
        import pandas

        from .diagnostics import ess_bulk, rhat

        rows = {}
        for var in (var_names if var_names is not None else list(data)):
            v = np.asarray(data[var], dtype=np.float64)          # [chain, iter(, k)]
            cols = {var: v} if v.ndim == 2 else {f"{var}[{i}]": v[..., i] for i in range(v.shape[-1])}
            for nm, x in cols.items():
                e = ess_bulk(x) if x.shape[1] >= 4 else float("nan")
                rows[nm] = {"mean": float(x.mean()), "sd": float(x.std(ddof=1)), "ess_bulk": e,
                            "mcse_mean": float(x.std(ddof=1) / np.sqrt(e)) if e == e and e > 0 else float("nan"),
                            "r_hat": rhat(x) if x.shape[1] >= 4 else float("nan")}
        return pandas.DataFrame.from_dict(rows, orient="index")

    arviz.summary = summary
    return {"arviz": arviz}


class _Unavailable(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        raise ImportError(f"{self.__name__} is not installed: plotting helpers are unavailable")


def install_import_shims(force=()):
    """Register the stand-ins for the packages that cannot be imported (or are listed in `force`).  Returns the
    names that were shimmed."""
    done = []
    if "jax" in force or _missing("jax"):
        sys.modules.update(_make_jax())
        done.append("jax")
    if "arviz" in force or _missing("arviz"):
        sys.modules.update(_make_arviz())
        done.append("arviz")
    if "matplotlib" in force or _missing("matplotlib"):
        mpl = _Unavailable("matplotlib")
        plt = _Unavailable("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        mpl.__dict__["pyplot"] = plt
        done.append("matplotlib")
    return done
