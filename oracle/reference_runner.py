"""TEST INFRASTRUCTURE ONLY.  Runs the REFERENCE'S OWN `sde/mici_extensions.py` (unmodified, loaded from the read-only
checkout) in this container: `jax` is the torch-backed stand-in of `oracle/jax_torch_shim.py`, `mici` the minimal
stand-in of `manifold_mcmc_for_diffusions_b200.mici_compat` (Mici 0.1.10 is not installable here), and the model
callables are the reference's own `sde/example_models/*.py` executed through a SymPy-backed `symnum` stand-in
(`oracle/symnum_sympy_shim.py`).  Used to pin the
restated oracle (`tests/test_reference_pin_cpu.py`) and to generate `tests/golden/reference_pin_golden.npz`
(`tests/golden/make_golden_reference_pin.py`).  The GPU box has no reference checkout: it only sees the committed
vectors."""
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("MMD_REFERENCE_ROOT", "/root/reference")
_cached = {}


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "sde", "mici_extensions.py"))


def load():
    """The reference module object (its real source file executed with the stand-in jax / mici)."""
    if "mod" in _cached:
        return _cached["mod"]
    from oracle.jax_torch_shim import install

    install()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from manifold_mcmc_for_diffusions_b200 import mici_compat

    saved = {k: sys.modules.get(k) for k in ["mici"] + ["mici." + s for s in (
        "adapters", "errors", "integrators", "matrices", "samplers", "solvers", "states", "systems", "transitions")]}
    try:
        import mici  # noqa: F401  (a real Mici would be used if it were installed)
    except ImportError:
        sys.modules["mici"] = mici_compat
        for sub in ("adapters", "errors", "integrators", "matrices", "samplers", "solvers", "states", "systems",
                    "transitions"):
            sys.modules["mici." + sub] = getattr(mici_compat, sub)
    spec = importlib.util.spec_from_file_location("reference_sde_mici_extensions",
                                                  os.path.join(REF_ROOT, "sde", "mici_extensions.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # leave sys.modules as it was for "mici" (the product's alias installer decides that for itself)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    _cached["mod"] = mod
    return mod


def make_fhn_system(obs_interval, num_steps_per_obs, num_obs_per_subseq, y_seq, noise=0, sigma=0.1,
                    use_gaussian_splitting=False, models="reference"):
    """The reference's ConditionedDiffusionConstrainedSystem for the FHN model (noise: 0 noiseless, 1 fixed
    observation noise scale, 2 inferred scale).  models="reference": the model callables are the reference's own
    `sde/example_models/fhn.py` (SymNum stand-in); "oracle": the torch functions of oracle/models.py."""
    import torch

    if models == "reference":
        fhn = load_models()[0]
    else:
        from oracle.models import fhn

    ref = load()
    dim_u = 5 if noise == 2 else 4
    gen_sigma = None if noise == 0 else (float(sigma) if noise == 1 else fhn.generate_σ_y)
    return ref.ConditionedDiffusionConstrainedSystem(
        obs_interval, num_steps_per_obs, num_obs_per_subseq, torch.as_tensor(y_seq, dtype=torch.float64), dim_u, 2, 2,
        fhn.forward_func, fhn.generate_x_0, fhn.generate_z, fhn.obs_func, gen_sigma, use_gaussian_splitting,
        dim_v_0=2)


def load_models():
    """The reference's own model modules `sde/example_models/fhn.py` and `sir.py` (with `sde/integrators.py` and
    `sde/transforms.py`), executed unmodified: `symnum` is the SymPy stand-in of oracle/symnum_sympy_shim.py, `jax`
    the torch stand-in.  Returns (fhn, sir)."""
    if "models" in _cached:
        return _cached["models"]
    import types

    from oracle.jax_torch_shim import install
    from oracle.symnum_sympy_shim import install as install_symnum

    install()
    install_symnum()
    names = ["sde", "sde.integrators", "sde.transforms", "sde.example_models"]
    saved = {k: sys.modules.get(k) for k in names}

    def by_path(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, *rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    try:
        pkg = types.ModuleType("sde")
        sys.modules["sde"] = pkg
        pkg.integrators = sys.modules["sde.integrators"] = by_path("sde.integrators", ("sde", "integrators.py"))
        pkg.transforms = sys.modules["sde.transforms"] = by_path("sde.transforms", ("sde", "transforms.py"))
        fhn = by_path("reference_sde_example_models_fhn", ("sde", "example_models", "fhn.py"))
        sir = by_path("reference_sde_example_models_sir", ("sde", "example_models", "sir.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached["models"] = (fhn, sir)
    return fhn, sir
