"""Distributional known-answer test against the reference's own recorded output:
FitzHugh-Nagumo_example.ipynb fully specifies a posterior (data from RandomState(20200710), notebook
prior parametrisation, obs_interval 0.5, 100 x 25 steps, Gaussian splitting, Newton solver, partition
switching) and records its ArviZ summary (cell 45).  The on-device constrained HMC sampler must reproduce
the posterior means within Monte Carlo error and the posterior standard deviations within 15 %.
(profiles/r1_notebook_known_answer.json holds a longer run: 256 chains x 800 transitions, max |z| = 2.4.)"""

import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_posterior_matches_notebook_table():
    env = dict(os.environ, NCH="128", NBURN="800", NMAIN="500", L="8", DT="0.1", USCALE="1.0", SOLVER="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "notebook_known_answer.py")], env=env,
                         capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    res = json.loads(out)
    assert res["stuck_chain_fraction"] == 0.0
    assert 0.9 < res["accept_stat"] <= 1.0
    for name, v in res["vars"].items():
        assert v["rhat"] < 1.05, (name, v)
        assert abs(v["z"]) < 4.0, (name, v)                       # mean within MC error of the notebook's
        assert abs(v["sd"] / v["notebook_sd"] - 1.0) < 0.15, (name, v)
