"""TEST INFRASTRUCTURE ONLY.  Runs the REFERENCE'S OWN `sde/mici_extensions.py` (unmodified, loaded from the read-only
checkout) in this container: `jax` is the torch-backed stand-in of `oracle/jax_torch_shim.py`, `mici` the minimal
stand-in of `manifold_mcmc_for_diffusions_b200.mici_compat` (Mici 0.1.10 is not installable here), the model callables
are the torch functions of `oracle/models.py` (the reference's SymNum-generated ones need symnum).  Used to pin the
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
                    use_gaussian_splitting=False):
    """The reference's ConditionedDiffusionConstrainedSystem for the FHN model (noise: 0 noiseless, 1 fixed
    observation noise scale, 2 inferred scale)."""
    import torch

    from oracle.models import fhn

    ref = load()
    dim_u = 5 if noise == 2 else 4
    gen_sigma = None if noise == 0 else (float(sigma) if noise == 1 else fhn.generate_σ_y)
    return ref.ConditionedDiffusionConstrainedSystem(
        obs_interval, num_steps_per_obs, num_obs_per_subseq, torch.as_tensor(y_seq, dtype=torch.float64), dim_u, 2, 2,
        fhn.forward_func, fhn.generate_x_0, fhn.generate_z, fhn.obs_func, gen_sigma, use_gaussian_splitting,
        dim_v_0=2)
