"""Host-side (NumPy) mirrors of ``sde.example_models`` used either side of the CUDA hot path: data
simulation, trace functions, initial states.  Every callable carries a ``_mmd_model`` tag; the
constrained system recognises the tag and maps the model to its generated device functor -- arbitrary
Python callables cannot be traced into CUDA and are rejected, never run on a CPU fallback."""
from . import fhn, fhn_notebook, sir  # noqa: F401
