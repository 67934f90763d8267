"""The C-ABI shared library loads and exports every symbol include/mmd_b200.h declares (CPU only;
no compute calls), and refuses to run without a GPU instead of falling back to a CPU path."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g

    return g.build()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mmd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(pkg):
    from manifold_mcmc_for_diffusions_b200 import _lib

    L = pkg.lib()
    names = _declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in mmd_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_no_cpu_fallback(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from manifold_mcmc_for_diffusions_b200 import BatchedChains, MmdError

    with pytest.raises(MmdError, match="no CUDA device"):
        BatchedChains("fhn", 0.2, 5, 5, [[0.0]] * 10, 4, 2)


def test_product_does_not_import_oracle():
    pkgdir = os.path.join(ROOT, "manifold_mcmc_for_diffusions_b200")
    for dp, _, fs in os.walk(pkgdir):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_dlpack_consumer_rejects_host_memory():
    """The DLPack path is device-only: CPU producers are refused (no hidden host fallback)."""
    import numpy as np
    import pytest
    import torch

    from manifold_mcmc_for_diffusions_b200._dlpack import DeviceArray

    with pytest.raises(ValueError):
        DeviceArray(torch.zeros(3, dtype=torch.float64))
    with pytest.raises(ValueError):
        DeviceArray(np.zeros(3))
    with pytest.raises(TypeError):
        DeviceArray([0.0, 1.0])
