"""Minimal stand-in for the subset of Mici 0.1.10 that ``sde.mici_extensions`` and the CHMC scripts
import (the real package is not installable in this environment).

Only used when ``import mici`` fails: ``manifold_mcmc_for_diffusions_b200.mici_extensions`` prefers
the real Mici.  Semantics follow Mici 0.1.10 as recalled in SURVEY.md (state cache keyed by
``(type(system).__name__, id(system), method name)``, cache invalidation on variable assignment,
``ConstrainedLeapfrogIntegrator._step`` = A(dt/2) B(dt) A(dt/2) with reversibility check,
multinomial dynamic integration transition, dual-averaging step-size adapter)."""

from . import adapters, errors, integrators, matrices, samplers, solvers, states, systems, transitions  # noqa: F401
