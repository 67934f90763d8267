"""GPU parity on the reference's BUNDLED data sets at the defaults of its experiment scripts (BASELINE.json configs 2
and 3), against oracle-frozen vectors (tests/golden/make_golden_bundled.py, which read the .npz files from the
reference checkout; the observations travel inside the fixture):

  config 2  scripts/fhn_model_noisy_obs_chmc_experiment.py:61-64 -- fhn_model_noisy_obs_simulated_data.npz,
            T=100, S=40, R=5, fixed observation noise 0.1
  config 3  scripts/sir_model_chmc_experiment.py:62-79 -- sir_model_boarding_school_data.npz (14 daily counts
            3, 8, 28, 75, 221, 281, ...), S=20, one block of all observations, fixed observation noise 1.0

Tolerances: constraint 1e-11 absolute (scaled), log-det 1e-10, gradients / positions 1e-9 relative, momenta 1e-8,
identical projection iteration counts with both solvers."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    "fhn_noisy": dict(model="fhn", dim_u=4),
    "sir": dict(model="sir", dim_u=4),
}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "bundled_configs_golden.npz"))


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _make(gold, tag, n=1):
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    cfg = CONFIGS[tag]
    S = int(gold[f"{tag}_S"])
    R = int(gold[f"{tag}_R"]) if f"{tag}_R" in gold.files else int(gold[f"{tag}_T"])
    return BatchedChains(cfg["model"], float(gold[f"{tag}_obs_interval"]), S, R, gold[f"{tag}_y"], cfg["dim_u"], n,
                         noise=1, sigma_fixed=float(gold[f"{tag}_sigma"]))


def test_bundled_series_is_the_reference_one(gold):
    # first entries of scripts/sir_model_boarding_school_data.npz (the 1978 boarding-school influenza counts)
    assert gold["sir_y"][:6, 0].tolist() == [3.0, 8.0, 28.0, 75.0, 221.0, 281.0]
    assert int(gold["fhn_noisy_T"]) == 100 and int(gold["fhn_noisy_S"]) == 40 and int(gold["fhn_noisy_R"]) == 5


@pytest.mark.parametrize("tag", list(CONFIGS))
def test_point_quantities(gold, tag):
    bc = _make(gold, tag)
    q0, xo, p_raw = gold[f"{tag}_q0"][None], gold[f"{tag}_xobs"][None], gold[f"{tag}_p_raw"][None]
    bc.set_state(q0, xo, 0, p=p_raw)
    c_ref = gold[f"{tag}_c"]
    assert np.max(np.abs(bc.constr()[0] - c_ref)) < 1e-11 * max(1.0, np.max(np.abs(gold[f"{tag}_y"])))
    bc.linearize(True)
    ld = float(gold[f"{tag}_ld"])
    assert abs(bc.log_det_sqrt_gram()[0] - ld) < 1e-10 * max(1.0, abs(ld))
    assert _rel(bc.grad_log_det_sqrt_gram()[0], gold[f"{tag}_grad_ld"]) < 1e-9
    assert _rel(bc.normal_space_component(p_raw)[0], gold[f"{tag}_nsc"]) < 1e-9
    bc.close()


@pytest.mark.parametrize("solver", ["quasi_newton", "newton"])
@pytest.mark.parametrize("tag", list(CONFIGS))
def test_leapfrog_step(gold, tag, solver):
    bc = _make(gold, tag)
    bc.opts.solver = 1 if solver == "newton" else 0
    q0, xo, p_raw = gold[f"{tag}_q0"][None], gold[f"{tag}_xobs"][None], gold[f"{tag}_p_raw"][None]
    bc.set_state(q0, xo, 0, p=p_raw)
    bc.linearize(True)
    bc.project_momentum()
    bc.leapfrog_step(float(gold[f"{tag}_dt"]))
    info = bc.step_info()
    q, p, _ = bc.get_state()
    assert info["status"][0] == 0
    it = gold[f"{tag}_{solver}_it"]
    assert info["iters_fwd"][0] == it[0] and info["iters_rev"][0] == it[1]
    assert _rel(q[0], gold[f"{tag}_{solver}_q"]) < 1e-9
    assert _rel(p[0], gold[f"{tag}_{solver}_p"]) < 1e-8
    h = bc.hamiltonian()[0]
    h_ref = float(gold[f"{tag}_{solver}_h"])
    assert abs(h - h_ref) < 1e-9 * max(1.0, abs(h_ref))
    bc.close()
