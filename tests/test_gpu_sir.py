"""GPU parity for the SIR model (sde/example_models/sir.py: Euler-Maruyama on the log-transformed SDE,
non-linear observation exp(x[1]), non-diagonal generate_z, inferred observation noise scale) against the
float64 autodiff oracle: constraint, log-det, its gradient (incl. the observation-curvature term),
normal-space projection and constrained leapfrog steps with both projection solvers."""

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from oracle.models import sir

pytestmark = pytest.mark.gpu

CASES = [(6, 4, 6), (6, 4, 3), (7, 3, 3)]   # (T, S, R): single block, two blocks, ragged


def make_sir_problem(T, S, R, n_chains, seed=11):
    rng = np.random.default_rng(seed)
    obs_interval = 1.0
    delta = obs_interval / S
    qs, xs = [], []
    # data from one simulated path
    u_true = np.array([-1.0, -0.3, 0.7, 0.2, np.log(0.5)])   # recovery rate 0.37 / day, log contact rate ~ 0.7
    z = sir.generate_z(torch.tensor(u_true[:4]))
    x = sir.generate_x_0(z, torch.tensor([0.7]))
    ys = []
    for t in range(T * S):
        x = sir.forward_func(z, x, torch.tensor(0.3 * rng.standard_normal(3)), delta)
        if (t + 1) % S == 0:
            ys.append(float(torch.exp(x[1])))
    y = np.array(ys)[:, None] + 0.5 * rng.standard_normal((T, 1))
    system = O.OracleSystem(obs_interval, S, R, y, 5, 3, 3, sir.forward_func, sir.generate_x_0, sir.generate_z,
                            sir.obs_func, sir.generate_σ_y, False, dim_v_0=1)
    for c in range(n_chains):
        crng = np.random.default_rng([seed, c])
        u = u_true + 0.1 * crng.standard_normal(5)
        v0 = np.array([0.7 + 0.1 * crng.standard_normal()])
        v = 0.3 * crng.standard_normal((T * S, 3))
        zc = sir.generate_z(torch.tensor(u[:4]))
        sig = float(np.exp(u[4]))
        xx = sir.generate_x_0(zc, torch.tensor(v0))
        xobs, n = [], []
        for t in range(T * S):
            xx = sir.forward_func(zc, xx, torch.tensor(v[t]), delta)
            if (t + 1) % S == 0:
                xobs.append(xx.numpy().copy())
                n.append((y[(t + 1) // S - 1, 0] - float(torch.exp(xx[1]))) / sig)   # puts the state on the manifold
        qs.append(np.concatenate([u, v0, v.reshape(-1), np.array(n)]))
        xs.append(np.stack(xobs))
    assert np.min(np.stack(xs)[..., :2]) > -10.0   # a live epidemic: far from the -500 clip of sir.py:54-70
    return dict(T=T, S=S, R=R, y=y, system=system, q=np.stack(qs), xobs=np.stack(xs), obs_interval=obs_interval)


@pytest.fixture(scope="module", params=CASES, ids=lambda c: "T%d_S%d_R%d" % c)
def prob(request):
    T, S, R = request.param
    return make_sir_problem(T, S, R, n_chains=3)


def make_bc(prob):
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    return BatchedChains("sir", prob["obs_interval"], prob["S"], prob["R"], prob["y"], 5, prob["q"].shape[0], noise=2)


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _parts(prob):
    return range(prob["system"].num_partition)


def test_point_quantities(prob):
    sysm = prob["system"]
    rng = np.random.default_rng(1)
    for part in _parts(prob):
        for off in (0.0, 0.02):
            q = prob["q"] + off * rng.standard_normal(prob["q"].shape)
            bc = make_bc(prob)
            assert bc.dim_q == q.shape[1] and bc.num_partition == sysm.num_partition
            bc.set_state(q, prob["xobs"], part)
            c = bc.constr()
            bc.linearize(True)
            ld, g = bc.log_det_sqrt_gram(), bc.grad_log_det_sqrt_gram()
            vct = rng.standard_normal(q.shape)
            nsc = bc.normal_space_component(vct)
            bc.update_x_obs_seq()
            _, _, xg = bc.get_state()
            for i in range(q.shape[0]):
                c_o = sysm._constr(torch.tensor(q[i]), torch.tensor(prob["xobs"][i]), part).numpy()
                assert c.shape[1] == c_o.shape[0]
                assert np.max(np.abs(c[i] - c_o)) < 1e-11 * max(1.0, np.abs(c_o).max())
                if off == 0.0:
                    assert np.max(np.abs(c_o)) < 1e-10
                pt = sysm.point(q[i], prob["xobs"][i], part)
                assert abs(ld[i] - pt["ld"]) < 1e-10 * max(1.0, abs(pt["ld"]))
                assert _rel(g[i], pt["grad_ld"].numpy()) < 1e-9
                nsc_o = sysm._normal_space_component(torch.tensor(vct[i]), pt["jac"], pt["chol"]).numpy()
                assert _rel(nsc[i], nsc_o) < 1e-9
                assert np.max(np.abs(xg[i] - sysm._generate_x_obs_seq(torch.tensor(q[i])).numpy())) < 1e-12
            bc.close()


@pytest.mark.parametrize("solver", ["quasi_newton", "newton"])
def test_leapfrog_steps(prob, solver):
    sysm = prob["system"]
    q0, xo = prob["q"], prob["xobs"]
    dt = 0.02
    rng = np.random.default_rng(3)
    p_raw = rng.standard_normal(q0.shape)
    for part in _parts(prob):
        bc = make_bc(prob)
        bc.opts.solver = 1 if solver == "newton" else 0
        bc.set_state(q0, xo, part, p=p_raw)
        bc.linearize(True)
        bc.project_momentum()
        traj = []
        for s in range(2):
            bc.leapfrog_step(dt)
            info = bc.step_info()
            qg, pg, _ = bc.get_state()
            traj.append((qg, pg, info, bc.hamiltonian()))
        for i in range(q0.shape[0]):
            pt = sysm.point(q0[i], xo[i], part)
            p = sysm.project_onto_cotangent_space(torch.tensor(p_raw[i]), pt)
            q = torch.tensor(q0[i])
            for s in range(2):
                q, p, pt, inf = O.leapfrog_step(sysm, q, p, xo[i], part, dt, pt=pt, solver=solver)
                qg, pg, info, hg = traj[s]
                assert info["status"][i] == 0
                assert info["iters_fwd"][i] == inf["n_fwd"] and info["iters_rev"][i] == inf["n_back"]
                assert _rel(qg[i], q.numpy()) < 1e-9
                assert _rel(pg[i], p.numpy()) < 1e-8
                assert abs(hg[i] - sysm.h(q, p, pt)) < 1e-9 * abs(hg[i])
        bc.close()


def test_sir_script_shape_golden():
    """The reference's SIR experiment shape (14 observations x 20 steps, one block, inferred noise scale,
    dim_q = 860; a synthetic epidemic curve of the boarding-school shape) against committed oracle-frozen vectors (tests/golden/make_golden_sir.py)."""
    import os

    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sir_T14_S20_golden.npz"))
    n = g["q0"].shape[0]
    for solver, sid in (("quasi_newton", 0), ("newton", 1)):
        bc = BatchedChains("sir", 1.0, int(g["S"]), int(g["T"]), g["y"], 5, n, noise=2)
        assert bc.dim_q == 860 and bc.num_partition == 1
        bc.opts.solver = sid
        bc.set_state(g["q0"], g["xobs"], 0, p=g["p_raw"])
        assert np.max(np.abs(bc.constr() - g["c"])) < 1e-9
        bc.linearize(True)
        assert np.max(np.abs(bc.log_det_sqrt_gram() - g["ld"])) < 1e-9 * np.max(np.abs(g["ld"]))
        assert _rel(bc.grad_log_det_sqrt_gram(), g["grad_ld"]) < 1e-9
        assert _rel(bc.normal_space_component(g["p_raw"]), g["nsc"]) < 1e-9
        bc.project_momentum()
        bc.leapfrog_step(float(g["dt"]))
        info = bc.step_info()
        q, p, _ = bc.get_state()
        assert np.all(info["status"] == 0)
        assert np.array_equal(info["iters_fwd"], g[f"{solver}_it"][:, 0])
        assert np.array_equal(info["iters_rev"], g[f"{solver}_it"][:, 1])
        assert _rel(q, g[f"{solver}_q"]) < 1e-9
        assert _rel(p, g[f"{solver}_p"]) < 1e-8
        h = bc.hamiltonian()
        assert np.max(np.abs(h - g[f"{solver}_h"]) / np.abs(h)) < 1e-9
        bc.close()


def test_clip_region_derivatives():
    """sir.py:54-70: a log-state component at or below -500 is frozen at -500 and, as autodiff sees the clip and the
    select, has no sensitivity and exerts none.  Blocks that start from a conditioned state with log S = -600 run
    their whole interval inside the clip region; constraint, log-det, its gradient and the normal-space component
    must agree with the autodiff oracle there too."""
    prob = make_sir_problem(6, 4, 3, n_chains=2, seed=17)
    sysm = prob["system"]
    xo = prob["xobs"].copy()
    xo[:, :, 0] = -600.0
    rng = np.random.default_rng(5)
    # partition 0: blocks [0:3] (live, conditioned on its full end state) and [3:6] (starts held, observation rows
    # only).  In the shifted partition an interior block would start held AND be conditioned on its full end state:
    # that constraint row has no gradient and the Gram matrix is singular for the reference as well.
    for part in (0,):
        q = prob["q"]
        bc = make_bc(prob)
        bc.set_state(q, xo, part)
        c = bc.constr()
        bc.linearize(True)
        ld, g = bc.log_det_sqrt_gram(), bc.grad_log_det_sqrt_gram()
        vct = rng.standard_normal(q.shape)
        nsc = bc.normal_space_component(vct)
        for i in range(q.shape[0]):
            c_o = sysm._constr(torch.tensor(q[i]), torch.tensor(xo[i]), part).numpy()
            assert np.max(np.abs(c[i] - c_o)) < 1e-11 * max(1.0, np.abs(c_o).max())
            pt = sysm.point(q[i], xo[i], part)
            assert np.isfinite(pt["grad_ld"].numpy()).all()
            assert abs(ld[i] - pt["ld"]) < 1e-10 * max(1.0, abs(pt["ld"]))
            assert _rel(g[i], pt["grad_ld"].numpy()) < 1e-9
            nsc_o = sysm._normal_space_component(torch.tensor(vct[i]), pt["jac"], pt["chol"]).numpy()
            assert _rel(nsc[i], nsc_o) < 1e-9
        bc.close()
