#!/usr/bin/env python
"""Runs one of the reference's scripts UNCHANGED on top of this package (needs a GPU and a checkout of the reference):

    python tools/run_reference_script.py /path/to/reference/scripts/fhn_model_noiseless_obs_chmc_experiment.py \
        --num-chain 2 --num-warm-up-iter 20 --num-main-iter 20 --projection-solver quasi-newton

`import sde`, `import mici` resolve to this package (install_reference_aliases) and `jax`, `arviz`, `matplotlib` to
the stand-ins of compat_shims when they are not installed; the script's own directory goes on sys.path so that its
`from utils import ...` works; everything else is the script's own code."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import install_reference_aliases  # noqa: E402
from manifold_mcmc_for_diffusions_b200.compat_shims import install_import_shims  # noqa: E402

if len(sys.argv) < 2:
    raise SystemExit(__doc__)
script = os.path.abspath(sys.argv[1])
install_reference_aliases()
print("import stand-ins:", install_import_shims(), file=sys.stderr)
sys.path.insert(0, os.path.dirname(script))
sys.argv = [script] + sys.argv[2:]
runpy.run_path(script, run_name="__main__")
