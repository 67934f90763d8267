#!/usr/bin/env python
"""Headline benchmark: constrained-HMC leapfrog steps/s over all chains, FHN noiseless T=100, S=25, R=5.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one ConstrainedLeapfrogIntegrator.step (both projection solves, reversibility check,
grad log-det) for every chain of the rank (default 4096 chains per GPU, weak scaling).  Momentum
refresh (Philox, on device), the Metropolis accept and the partition switch happen every
`--traj-len` steps INSIDE the timed region; only successful chain-steps are counted.
Prints ONE JSON line (see DESIGN.md "Measurement").
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T, S, R = 100, 25, 5
OBS_INTERVAL = 0.2
SEED = 20200710
METRIC = "chmc_leapfrog_steps_per_s"
UNIT = "chain-steps/s"
WORKLOAD = "FHN noiseless-obs CHMC, T=100 obs x S=25 steps/obs, R=5, quasi-Newton projection, float64"


def load_y():
    return np.load(os.path.join(ROOT, "tests", "golden", "fhn_yseq_T100.npy"))


def init_inputs(n, rank):
    """Synthetic initial states: the reference recipe (fhn_model_noiseless_obs_chmc_experiment.py:
    120-134) with a per-rank generator."""
    y = load_y()
    rng = np.random.default_rng([SEED, rank])
    u = rng.standard_normal((n, 4))
    v0 = rng.standard_normal((n, 2))
    xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
    return y, u, v0, xo


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip().split(",")
                self.samples.append((float(out[0]), float(out[1])))
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons)}


# work model (SURVEY.md 8d / DESIGN.md): bytes one quasi-Newton sweep must read per chain
def leapfrog_bytes_per_chain_step(qn_iters, newton=False):
    """Algorithmic HBM bytes of one constrained leapfrog step of one chain: doubles
    per SDE time step summed over the phases (X=V=2: a K_t record is 4 doubles, v_t / p_t / x_t records 2)."""
    # momentum projections (pass 1: J p, pass 2: p - J^T lambda), h1 kick and h2_flow fused in:
    project = (4 + 2 + 2 + 2 + 2) + (4 + 2 + 2 + 2 + 2)   # A(dt/2) at the old point + flow -> qw, pw
    project += (4 + 2) + (4 + 2 + 2 + 2 + 2)              # tangent projection at the new point + back flow
    project += (4 + 2 + 2 + 2 + 2) + (4 + 2 + 2)          # A(dt/2) at the new point
    # quasi-Newton: one sweep per iteration (K_t + work position); Newton: forward sweep that also stores the
    # trajectory (4 + 2 + 2) and a backward sweep re-linearising at the iterate (trajectory, work position, K_t)
    qn = (16 if newton else 6) * qn_iters
    qn_final = (4 + 2 + 2 + 2 + 2) + (4 + 2 + 2)           # forward: write q, p; reverse: compare with q_prev
    point = 4 + 8 + 12 + 10                                # trajectory, compressed Jacobian, tangent, adjoint sweeps
    return 8 * T * S * (project + qn + qn_final + point)


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from manifold_mcmc_for_diffusions_b200 import BatchedChains

    n = args.chains
    y, u, v0, xo = init_inputs(n, rank)
    bc = BatchedChains("fhn", OBS_INTERVAL, S, R, y, 4, n, device=local_rank)
    bc.set_chain_offset(rank * n)   # disjoint Philox streams per rank
    bc.opts.solver = 1 if args.solver == "newton" else 0
    if args.regroup:
        bc.set_chain_regrouping(True)   # per-chain results unchanged; chains with similar iteration counts share tiles
    bc.init_linear_interpolation(u, v0, xo, 0)
    L = args.traj_len
    # untimed burn-in towards the typical set (the linear-interpolation states are far in the tails)
    for it in range(args.burnin):
        bc.hmc_transition(args.burnin_dt, L, SEED + rank, it)
    bc.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        bc.synchronize()

    it_counter = [args.burnin]

    def do_steps(k):
        done = 0
        while done < k:
            m = min(L, k - done)
            bc.transition_begin(SEED + rank, it_counter[0])
            bc.transition_steps(args.dt, m)          # m leapfrog steps, one fused launch
            bc.transition_end(SEED + rank, it_counter[0], True)
            it_counter[0] += 1
            done += m

    do_steps(args.warmup)
    bc.successful_steps(reset=True)
    bc.total_qn_iterations(reset=True)
    launches0 = bc.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    bc.profile_enable(True, 64 * args.steps + 64)
    barrier()
    bc.timer_start()
    do_steps(args.steps)
    ms = bc.timer_stop_ms()
    barrier()
    sampler.stop_flag = True
    bc.profile_enable(False)
    launches = bc.launch_count() - launches0
    ok = bc.successful_steps()
    qn_iters = bc.total_qn_iterations()
    st = bc.transition_stats()
    info = bc.step_info()
    n_lf, ms_lf = bc.profile_summary(3)

    # ---- end to end through the public API with host buffers (pinned), every step ----
    # The chains are split over `e2e_pipeline` BatchedChains objects (each with its own stream): while one part
    # computes, the next part's inputs are uploaded (set_state(blocking=False)).  Every step still uploads q, p and
    # x_obs_seq of EVERY chain from pinned host memory and reads every chain's status / iteration counts back.
    e2e_steps = min(args.steps, args.e2e_steps)
    q_h, p_h, x_h = bc.get_state()
    part = bc.partition
    n_parts = max(1, args.e2e_pipeline)
    bounds = [n * i // n_parts for i in range(n_parts + 1)]
    if n_parts == 1:
        parts = [bc]
    else:
        parts = [BatchedChains("fhn", OBS_INTERVAL, S, R, y, 4, bounds[i + 1] - bounds[i], device=local_rank)
                 for i in range(n_parts)]
        for b in parts:
            b.opts.solver = bc.opts.solver
    pins = []
    for i in range(n_parts):
        sl = slice(bounds[i], bounds[i + 1])
        pins.append([torch.from_numpy(np.ascontiguousarray(a[sl])).pin_memory().numpy() for a in (q_h, p_h, x_h)])

    def issue(i):
        qn_, pn_, xn_ = pins[i]
        parts[i].set_state(qn_, xn_, part, p=pn_, blocking=False)   # H2D of this step's inputs (async, own stream)
        parts[i].leapfrog_step(args.dt)

    def collect(i):
        inf = parts[i].step_info()                # D2H of the step's result (status, iterations, reverse dist)
        return int((inf["status"] == 0).sum())

    def e2e_loop(k):
        # software pipeline over the parts: as soon as a part's results are back its next upload + step are queued,
        # so its copy runs while the other part computes
        good = 0
        for i in range(n_parts):
            issue(i)
        for s in range(k):
            for i in range(n_parts):
                good += collect(i)
                if s + 1 < k:
                    issue(i)
        return good

    e2e_loop(1)                                   # untimed: first touch of the staging buffers
    barrier()
    t0 = time.perf_counter()
    ok_e2e = e2e_loop(e2e_steps)
    for b in parts:
        b.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(sum(a.nbytes for pin in pins for a in pin))
    d2h = int(n * (4 + 4 + 4 + 8))
    if n_parts > 1:
        for b in parts:
            b.close()

    vals = torch.tensor([ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(ok), float(ok_e2e), float(launches)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = vals.tolist()
    ok_tot, ok_e2e_tot, launches_tot = tot.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # roofline of the dominant kernel (k_leapfrog: every phase of the step fused in one launch).
        # Algorithmic bytes: doubles moved per SDE time step and chain, summed over the
        # phases of one leapfrog step, with the MEASURED number of quasi-Newton sweeps.
        iters_per_step = qn_iters / max(ok, 1)          # forward + reverse iterations per successful step
        alg_bytes_step = leapfrog_bytes_per_chain_step(iters_per_step, newton=args.solver == "newton")
        alg_bytes = alg_bytes_step * ok                   # over the timed region, this rank
        lf_ms = ms_lf
        achieved = alg_bytes / (lf_ms * 1e-3) / 1e9 if lf_ms > 0 else 0.0
        out = {
            "metric": METRIC,
            "value": ok_tot / (ms_max * 1e-3),
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD if args.solver == "quasi-newton" else WORKLOAD.replace("quasi-Newton", "Newton"),
                "chains_per_gpu": n,
                "chains_per_cta_tile": bc.chains_per_tile(),
                "chain_regrouping": bool(args.regroup),
                "step_size": args.dt,
                "traj_len": L,
                "burnin_transitions": args.burnin,
                "l2": "per-step working set %.1f GB per GPU >> 126 MB L2 (inputs larger than L2)"
                % (n * 450e3 / 1e9),
                "step_success_frac": ok_tot / (n * world * args.steps),
                "accept_stat_last": float(st["accept_stat"].mean()),
                "qn_iters_per_step": iters_per_step,
            },
            "roofline": {
                "bound": "hbm",
                "kernel": "k_leapfrog (fused constrained leapfrog step: projections, quasi-Newton solves, linearise + grad log-det)",
                "achieved": achieved,
                "peak": hbm_peak,
                "unit": "GB/s",
                "frac": achieved / hbm_peak,
                "traffic": None,
                "peak_source": peak_src,
                "alg_bytes_per_chain_step": alg_bytes_step,
                "kernel_ms_per_launch": lf_ms / max(n_lf, 1),
                "launches": n_lf,
                "share_of_timed_region": lf_ms / ms,
            },
            "e2e": {
                "value": ok_e2e_tot / (e2e_ms_max * 1e-3),
                "unit": UNIT,
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "pipeline": "%d BatchedChains objects of %d chains, upload of one overlapped with compute of the other"
                % (n_parts, n // n_parts) if n_parts > 1 else "none",
            },
            "gpu_launches": int(launches_tot),
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.cpu_seconds, dt=args.dt, L=L, burnin=args.burnin,
                                               burnin_dt=args.burnin_dt)
        emit(out)
    bc.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# CPU baseline: the compiled (numba) restatement of the reference path on the host cores, bounded sample.
# Same workload as the GPU arm: same data, same initialisation recipe, same burn-in (transitions x trajectory
# length x burn-in step size), same step size, trajectory length, solver and tolerances, momentum refresh,
# Metropolis accept and partition switch inside the timed region; only successful chain-steps are counted.
# ---------------------------------------------------------------------------------------------
PUBLISHED_STEPS_PER_S_PER_CHAIN = 71.6   # FitzHugh-Nagumo_example.ipynb raw lines 716, 752 (2.53 it/s x 28.3 steps/it)


def _cpu_worker(args):
    chain, seconds, dt, L, burnin, burnin_dt = args
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["NUMBA_NUM_THREADS"] = "1"
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import numba_chmc as N

    y = load_y()
    rng = np.random.default_rng([SEED, 10_000 + chain])
    ch = N.NumbaChain(T, S, R, y, OBS_INTERVAL)
    q, xo = N.linear_interpolation_init(T, S, y, OBS_INTERVAL, rng)     # u, v_0 ~ N(0, I) like init_inputs()
    ch.set_state(q, xo, 0)
    for _ in range(burnin):
        ch.hmc_transition(burnin_dt, L, rng)
    ch.n_steps_ok = 0
    t0 = time.perf_counter()
    ntr = 0
    while time.perf_counter() - t0 < seconds:
        ch.hmc_transition(dt, L, rng)
        ntr += 1
    return ch.n_steps_ok, time.perf_counter() - t0, ntr


def cpu_baseline(seconds=15.0, dt=0.1, L=8, burnin=60, burnin_dt=0.05):
    import multiprocessing as mp
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import numba_chmc as N   # compile once here: the workers then load numba's on-disk cache

    y = load_y()
    ch = N.NumbaChain(T, S, R, y, OBS_INTERVAL)
    q, xo = N.linear_interpolation_init(T, S, y, OBS_INTERVAL, np.random.default_rng(0))
    ch.set_state(q, xo, 0)
    ch.hmc_transition(burnin_dt, 1, np.random.default_rng(0))
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(c, seconds, dt, L, burnin, burnin_dt) for c in range(cores)])
    wall = time.perf_counter() - t0
    rate = sum(r[0] / r[1] for r in res)            # chains run concurrently, one per core
    return {
        "value": rate,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "same_config": True,
        "per_core": rate / cores,
        "reference_published_per_chain": PUBLISHED_STEPS_PER_S_PER_CHAIN,
        "sample": f"{cores} chains (one process per core), each: linear-interpolation init, {burnin} untimed burn-in "
                  f"transitions of {L} steps at dt={burnin_dt}, then {sum(r[2] for r in res)} timed transitions of {L} "
                  f"steps at dt={dt} ({seconds:.0f} s per core; {wall:.0f} s wall incl. burn-in); numba-compiled "
                  f"float64 restatement oracle/numba_chmc.py (quasi-Newton), successful chain-steps only",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.solver != "quasi-newton":
        raise SystemExit("the compiled CPU restatement implements the quasi-Newton solver only")
    cb = cpu_baseline(seconds=max(2.0, min(60.0, 1.5 * args.steps)), dt=args.dt, L=args.traj_len, burnin=args.burnin,
                      burnin_dt=args.burnin_dt)
    out = {
        "impl": "reference",
        "metric": METRIC,
        "value": cb["value"],
        "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": (1e3 / cb["value"]) if cb["value"] > 0 else None,   # one chain-step on the host cores
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "step_size": args.dt, "traj_len": args.traj_len,
                   "burnin_transitions": args.burnin},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded below write there too (NCCL prints its version
    banner to stdout at NCCL_DEBUG=VERSION / WARN): from here on file descriptor 1 goes to stderr, and only
    `emit` writes to the real stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=16384, help="chains per GPU")
    ap.add_argument("--dt", type=float, default=0.1)
    ap.add_argument("--traj-len", type=int, default=8)
    ap.add_argument("--burnin", type=int, default=60)
    ap.add_argument("--burnin-dt", type=float, default=0.05)
    ap.add_argument("--solver", choices=("quasi-newton", "newton"), default="quasi-newton",
                    help="projection solver (north-star item 3 names the quasi-Newton loop; the reference scripts "
                         "default to Newton, scripts/utils.py:137-142)")
    ap.add_argument("--regroup", type=int, default=0,
                    help="re-assign chains to CTA tiles by iteration count at every partition switch (results per "
                         "chain are bit-identical, tests/test_gpu_regrouping.py; measured: no gain, the iteration "
                         "count of a chain is not persistent across transitions)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-pipeline", type=int, default=2,
                    help="number of BatchedChains objects the end-to-end loop alternates between (1 = no overlap)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
