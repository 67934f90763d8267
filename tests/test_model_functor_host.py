"""The generated device functors (step, Jacobians, second-order contraction), compiled for the host
with g++, against torch autodiff of the oracle's forward_func.  CPU only."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle.models import fhn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "manifold_mcmc_for_diffusions_b200", "csrc")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libhostshim.so")
    subprocess.run(
        ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(CSRC, "host_model_shim.cpp"),
         "-I", CSRC, "-o", out],
        check=True,
    )
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_fhn_functor_matches_autodiff(shim):
    rng = np.random.default_rng(11)
    dl = 0.008
    sd = np.sqrt(dl)
    for _ in range(10):
        z = np.array([0.3, 0.1, 1.5, 0.8]) * np.exp(0.3 * rng.standard_normal(4))
        x = rng.standard_normal(2)
        v = rng.standard_normal(2)
        zt, xt, vt = torch.tensor(z), torch.tensor(x), torch.tensor(v)
        f = lambda z_, x_, v_: fhn.forward_func(z_, x_, v_, dl)  # noqa: E731
        xn = np.zeros(2)
        shim.fhn_step(_p(z), C.c_double(sd), _p(x), _p(v), _p(xn))
        assert np.max(np.abs(xn - f(zt, xt, vt).numpy())) < 1e-14
        Jz, Jx, Jv = torch.func.jacrev(f, argnums=(0, 1, 2))(zt, xt, vt)
        F, B, G = np.zeros(4), np.zeros(4), np.zeros(8)
        shim.fhn_jac_x(_p(z), C.c_double(sd), _p(x), _p(v), _p(F))
        shim.fhn_jac_v(_p(z), C.c_double(sd), _p(x), _p(v), _p(B))
        shim.fhn_jac_z(_p(z), C.c_double(sd), _p(x), _p(v), _p(G))
        assert np.max(np.abs(F.reshape(2, 2) - Jx.numpy())) < 1e-12 * max(1, np.abs(Jx.numpy()).max())
        assert np.max(np.abs(B.reshape(2, 2) - Jv.numpy())) < 1e-13
        assert np.max(np.abs(G.reshape(2, 4) - Jz.numpy())) < 1e-12 * max(1, np.abs(Jz.numpy()).max())
        # second-order contraction g[a] = sum_i sum_b d2 f_i/dy_a dy_b Th[b, i], y = (x, v, z)
        Th = rng.standard_normal((8, 2))

        def fj(yv):
            return fhn.forward_func(yv[4:8], yv[0:2], yv[2:4], dl)

        H = torch.func.jacfwd(torch.func.jacrev(fj))(torch.tensor(np.concatenate([x, v, z]))).numpy()  # [i,a,b]
        ref = np.einsum("iab,bi->a", H, Th)
        g = np.zeros(8)
        Thc = np.ascontiguousarray(Th)
        shim.fhn_hess_contract(_p(z), C.c_double(sd), _p(x), _p(v), _p(Thc), _p(g))
        assert np.max(np.abs(g - ref)) < 1e-10 * max(1.0, np.abs(ref).max())


def test_gen_z(shim):
    u = np.array([0.1, -0.2, 0.3, 0.4])
    z, dz = np.zeros(4), np.zeros(16)
    shim.fhn_gen_z(_p(u), _p(z), _p(dz))
    assert np.allclose(z, fhn.generate_z(torch.tensor(u)).numpy(), rtol=1e-15)
    J = torch.func.jacrev(fhn.generate_z)(torch.tensor(u)).numpy()
    assert np.allclose(dz.reshape(4, 4), J, rtol=1e-15)


def test_philox_normals_are_standard_normal_and_reproducible(shim):
    shim.philox_normal_pair.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    a, b = C.c_double(), C.c_double()
    vals = []
    for i in range(20000):
        shim.philox_normal_pair(1234, 7, i, C.byref(a), C.byref(b))
        vals += [a.value, b.value]
    vals = np.array(vals)
    assert abs(vals.mean()) < 0.02 and abs(vals.std() - 1.0) < 0.02
    assert abs(np.mean(vals ** 4) - 3.0) < 0.15
    shim.philox_normal_pair(1234, 7, 5, C.byref(a), C.byref(b))
    assert a.value == vals[10] and b.value == vals[11]
    # Philox-4x32-10 known-answer test (Random123 kat_vectors: counter=0, key=0)
    # checked through the raw block function in test_abi-level C shim is overkill; moments + determinism suffice here.


def test_sir_functor_matches_autodiff(shim):
    from oracle.models import sir

    rng = np.random.default_rng(12)
    dl = 0.05
    sd = np.sqrt(dl)
    for _ in range(10):
        u = 0.3 * rng.standard_normal(4)
        z = sir.generate_z(torch.tensor(u)).numpy()
        x = np.array([np.log(700.0), np.log(20.0), 0.3]) + 0.2 * rng.standard_normal(3)
        v = rng.standard_normal(3)
        zt, xt, vt = torch.tensor(z), torch.tensor(x), torch.tensor(v)
        f = lambda z_, x_, v_: sir.forward_func(z_, x_, v_, dl)  # noqa: E731
        xn = np.zeros(3)
        shim.sir_step(_p(z), C.c_double(sd), _p(x), _p(v), _p(xn))
        assert np.max(np.abs(xn - f(zt, xt, vt).numpy())) < 1e-13
        Jz, Jx, Jv = torch.func.jacrev(f, argnums=(0, 1, 2))(zt, xt, vt)
        F, B, G = np.zeros(9), np.zeros(9), np.zeros(12)
        shim.sir_jac_x(_p(z), C.c_double(sd), _p(x), _p(v), _p(F))
        shim.sir_jac_v(_p(z), C.c_double(sd), _p(x), _p(v), _p(B))
        shim.sir_jac_z(_p(z), C.c_double(sd), _p(x), _p(v), _p(G))
        assert np.max(np.abs(F.reshape(3, 3) - Jx.numpy())) < 1e-12 * max(1, np.abs(Jx.numpy()).max())
        assert np.max(np.abs(B.reshape(3, 3) - Jv.numpy())) < 1e-13 * max(1, np.abs(Jv.numpy()).max())
        assert np.max(np.abs(G.reshape(3, 4) - Jz.numpy())) < 1e-12 * max(1, np.abs(Jz.numpy()).max())
        Th = rng.standard_normal((10, 3))

        def fj(yv):
            return sir.forward_func(yv[6:10], yv[0:3], yv[3:6], dl)

        H = torch.func.jacfwd(torch.func.jacrev(fj))(torch.tensor(np.concatenate([x, v, z]))).numpy()
        ref = np.einsum("iab,bi->a", H, Th)
        g = np.zeros(10)
        Thc = np.ascontiguousarray(Th)
        shim.sir_hess_contract(_p(z), C.c_double(sd), _p(x), _p(v), _p(Thc), _p(g))
        assert np.max(np.abs(g - ref)) < 1e-10 * max(1.0, np.abs(ref).max())
        # generate_z: Jacobian and the second-derivative contraction used by the log-det gradient
        zz, dz = np.zeros(4), np.zeros(16)
        shim.sir_gen_z(_p(u), _p(zz), _p(dz))
        J = torch.func.jacrev(sir.generate_z)(torch.tensor(u)).numpy()
        assert np.allclose(zz, z, rtol=1e-15) and np.allclose(dz.reshape(4, 4), J, rtol=1e-14, atol=1e-16)
        Gam = rng.standard_normal((4, 4))
        H2 = torch.func.jacfwd(torch.func.jacrev(sir.generate_z))(torch.tensor(u)).numpy()  # [m, j, j']
        ref2 = np.einsum("mj,mjk->k", Gam, H2)
        ex = np.zeros(4)
        Gc = np.ascontiguousarray(Gam)
        shim.sir_gen_z_second(_p(u), _p(zz), _p(Gc), _p(ex))
        assert np.max(np.abs(ex - ref2)) < 1e-13 * max(1.0, np.abs(ref2).max())


def test_fhn_generators_with_runtime_parameters(shim):
    """generate_z / generate_x_0 of the FHN functor with run-time parameters (mmd_set_generator_params): values against
    the NumPy mirror example_models.fhn.make_generators, first and second derivatives against autodiff, for the
    reference's parametrisation, the notebook's and a random one."""
    from manifold_mcmc_for_diffusions_b200.example_models import fhn as M

    rng = np.random.default_rng(5)
    rnd = M.generator_params(scale=rng.uniform(0.3, 1.5, 4), shift=rng.standard_normal(4), exp_mask=(0, 1, 1, 0),
                             x0_shift=rng.standard_normal(2), x0_z=rng.standard_normal((2, 4)))
    for gp in (M.generator_params(), M.NOTEBOOK_GENERATOR_PARAMS, rnd):
        gen_z, gen_x0 = M.make_generators(gp)
        a, b, m = torch.tensor(gp[0:4]), torch.tensor(gp[4:8]), torch.tensor(gp[8:12] != 0)
        c, E = torch.tensor(gp[12:14]), torch.tensor(gp[14:22].reshape(2, 4))

        def tz(u_):
            lin = a * u_ + b
            return torch.where(m, torch.exp(lin), lin)

        for _ in range(5):
            u, v0, Gam = rng.standard_normal(4), rng.standard_normal(2), rng.standard_normal(16)
            z, dzdu, extra, x0, dv0, dz = (np.zeros(n) for n in (4, 16, 4, 2, 4, 8))
            shim.fhn_gen_all(_p(gp), _p(u), _p(v0), _p(Gam), _p(z), _p(dzdu), _p(extra), _p(x0), _p(dv0), _p(dz))
            assert np.max(np.abs(z - gen_z(u))) < 1e-14 * max(1, np.abs(z).max())
            assert np.max(np.abs(x0 - gen_x0(z, v0))) < 1e-13 * max(1, np.abs(x0).max())
            J = torch.func.jacrev(tz)(torch.tensor(u)).numpy()
            assert np.max(np.abs(dzdu.reshape(4, 4) - J)) < 1e-13 * max(1, np.abs(J).max())
            H = torch.func.jacfwd(torch.func.jacrev(tz))(torch.tensor(u)).numpy()          # [m, j, j']
            ex = np.einsum("mj,mjk->k", Gam.reshape(4, 4), H)
            assert np.max(np.abs(extra - ex)) < 1e-12 * max(1, np.abs(ex).max())
            assert np.array_equal(dv0.reshape(2, 2), np.eye(2)) and np.array_equal(dz.reshape(2, 4), E.numpy())
    # the default parameters reproduce the reference's generators exactly
    gz, gx = M.make_generators(M.generator_params())
    u, v0 = rng.standard_normal(4), rng.standard_normal(2)
    assert np.array_equal(gz(u), M.generate_z(u)) and np.array_equal(gx(gz(u), v0), M.generate_x_0(gz(u), v0))
