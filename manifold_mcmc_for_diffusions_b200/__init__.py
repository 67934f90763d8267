"""B200-native constrained-HMC hot path of ``sde.mici_extensions``."""

import sys

from ._lib import MmdError, lib  # noqa: F401
from .batched import BatchedChains  # noqa: F401


def install_reference_aliases(force_mici_compat=False):
    """Make the reference's import names resolve to this package, so scripts written against the
    reference (``import sde``, ``sde.mici_extensions``, ``sde.example_models.fhn``, ``import mici``) run
    unchanged: registers ``sde`` / ``sde.mici_extensions`` / ``sde.example_models`` in ``sys.modules`` and, when
    the real Mici is not installed (or ``force_mici_compat``), the minimal ``mici`` stand-in."""
    import types

    from . import example_models, mici_extensions

    sde = types.ModuleType("sde")
    sde.mici_extensions = mici_extensions
    sde.example_models = example_models
    sys.modules["sde"] = sde
    sys.modules["sde.mici_extensions"] = mici_extensions
    sys.modules["sde.example_models"] = example_models
    sys.modules["sde.example_models.fhn"] = example_models.fhn
    sys.modules["sde.example_models.sir"] = example_models.sir
    have_mici = False
    if not force_mici_compat:
        try:
            import mici  # noqa: F401

            have_mici = True
        except ImportError:
            pass
    if not have_mici:
        from . import mici_compat

        sys.modules["mici"] = mici_compat
        for sub in ("adapters", "errors", "integrators", "matrices", "samplers", "solvers", "states", "systems",
                    "transitions"):
            sys.modules["mici." + sub] = getattr(mici_compat, sub)
    return sde
