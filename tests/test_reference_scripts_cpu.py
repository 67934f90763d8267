"""The reference's own scripts against this package's import surface, as far as that goes without a GPU: with
`install_reference_aliases()` + `install_import_shims()` the module-level imports of scripts/utils.py resolve, its
argument parsers build, and the experiment script runs up to the first device call, where it stops with the
"no CPU path" error instead of falling back.  (The reference checkout only exists in the build container: skipped
elsewhere.  On a GPU box with a checkout, tools/run_reference_script.py runs the scripts to the end.)"""

import argparse
import importlib
import os
import subprocess
import sys

import pytest

REF = "/root/reference/scripts"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def test_reference_utils_imports_and_parsers():
    code = f"""
import sys, argparse
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {REF!r})
from manifold_mcmc_for_diffusions_b200 import install_reference_aliases
from manifold_mcmc_for_diffusions_b200.compat_shims import install_import_shims
install_reference_aliases(); shimmed = install_import_shims()
import utils
utils.setup_jax()
p = argparse.ArgumentParser()
utils.add_common_experiment_args(p, 25, 250, 1000)
utils.add_chmc_experiment_args(p, 5)
a = p.parse_args([])
assert a.projection_solver == "newton" and a.num_inner_h2_step == 1, a
import sde, mici
assert hasattr(sde.mici_extensions, "ConditionedDiffusionConstrainedSystem")
assert hasattr(mici.integrators, "ConstrainedLeapfrogIntegrator")
print("OK", shimmed)
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


def test_experiment_script_reaches_the_device_and_has_no_cpu_fallback(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the script would run to the end (see tools/run_reference_script.py)")
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"),
         os.path.join(REF, "fhn_model_noiseless_obs_chmc_experiment.py"), "--num-obs", "10", "--num-chain", "1",
         "--num-warm-up-iter", "2", "--num-main-iter", "2", "--output-root-dir", str(tmp_path)],
        capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CUDA device: this library has no CPU path" in out.stderr, out.stderr[-3000:]
