"""Host logic of the multi-GPU path on CPU: chain sharding, the all-gather of traces (gloo,
world_size 2), and the R-hat / ESS estimators on known-answer inputs."""

import os
import socket

import numpy as np
import pytest

from manifold_mcmc_for_diffusions_b200 import diagnostics as D
from manifold_mcmc_for_diffusions_b200.parallel import shard_range


def test_shard_range_partitions_everything():
    for n, w in [(65536, 8), (10, 3), (7, 8), (4096, 1)]:
        got = []
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            got += list(range(lo, hi))
        assert got == list(range(n))


def test_iid_draws_have_full_ess_and_unit_rhat():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((8, 2000))
    assert abs(D.rhat(x) - 1.0) < 0.01
    e = D.ess_bulk(x)
    assert 0.85 * x.size < e < 1.15 * x.size


def test_ar1_ess_matches_theory():
    rng = np.random.default_rng(1)
    phi = 0.9
    n = 20000
    x = np.zeros((4, n))
    eps = rng.standard_normal((4, n)) * np.sqrt(1 - phi ** 2)
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + eps[:, t]
    e = D.ess_bulk(x)
    theory = x.size * (1 - phi) / (1 + phi)
    assert 0.75 * theory < e < 1.3 * theory


def test_rhat_detects_unmixed_chains():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((4, 1000))
    x[0] += 3.0
    assert D.rhat(x) > 1.2


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from manifold_mcmc_for_diffusions_b200.parallel import allgather_chains, shard_range as sr

    n_total = 7
    lo, hi = sr(n_total, rank, world)
    full = np.arange(n_total * 5 * 2, dtype=np.float64).reshape(n_total, 5, 2)
    got = allgather_chains(full[lo:hi])
    q.put((rank, np.array_equal(got, full)))
    dist.destroy_process_group()


def test_allgather_chains_gloo_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_summary_json_and_trace_files(tmp_path):
    """Output format the reference's loaders expect (scripts/utils.py:368-381, 484-569)."""
    import glob
    import json
    import os

    from manifold_mcmc_for_diffusions_b200.diagnostics import save_and_print_summary

    rng = np.random.default_rng(0)
    traces = {"σ": rng.standard_normal((4, 200)) * 0.1 + 0.8, "x_0": rng.standard_normal((4, 200, 2))}
    out = save_and_print_summary(str(tmp_path), traces, ["σ", "x_0"], 12.0, 0.13, call_counts={"constr": 1234},
                                 verbose=False)
    loaded = json.load(open(os.path.join(tmp_path, "summary.json")))
    assert loaded == out
    assert set(loaded["mean"]) == {"σ", "x_0[0]", "x_0[1]"}
    assert abs(loaded["mean"]["σ"] - 0.8) < 0.02 and loaded["r_hat"]["σ"] < 1.05
    assert loaded["total_constr_calls"] == 1234 and loaded["final_integrator_step_size"] == 0.13
    assert len(glob.glob(os.path.join(tmp_path, "trace_*_σ.npy"))) == 4
    assert np.load(os.path.join(tmp_path, "trace_2_x_0.npy")).shape == (200, 2)
