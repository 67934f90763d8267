#!/usr/bin/env python
"""Headline benchmark: constrained-HMC leapfrog steps/s over all chains, FHN noiseless T=100, S=25, R=5.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one ConstrainedLeapfrogIntegrator.step (both projection solves, reversibility check,
grad log-det) for every chain of the rank.  Default workload = BASELINE.json config 5: a fixed population of
65,536 chains sharded over the ranks (strong scaling; `--scaling weak --chains N` keeps N chains per GPU), with the
reference scripts' traced scalars read back every transition, one NCCL all-gather of the traces at the end and
rank-normalised split-R-hat / bulk ESS / ESS per second in the `diagnostics` object of the line.  Momentum
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


def init_inputs_global(lo, hi):
    """The same recipe for a FIXED population of chains sharded over the ranks: generators are keyed by blocks of
    4096 GLOBAL chain indices, so chain c gets the same initial state whatever the number of ranks."""
    y = load_y()
    n, blk = hi - lo, 4096
    u, v0, xo = np.empty((n, 4)), np.empty((n, 2)), np.empty((n, T, 2))
    for b in range(lo // blk, (hi - 1) // blk + 1):
        g = np.random.default_rng([SEED, 7, b])
        fu, fv, fx = g.standard_normal((blk, 4)), g.standard_normal((blk, 2)), g.standard_normal((blk, T, 1))
        c0, c1 = max(lo, b * blk), min(hi, (b + 1) * blk)
        src, dst = slice(c0 - b * blk, c1 - b * blk), slice(c0 - lo, c1 - lo)
        u[dst], v0[dst] = fu[src], fv[src]
        xo[dst] = np.concatenate((np.broadcast_to(y, (c1 - c0, T, 1)), 0.5 * fx[src]), -1)
    return y, u, v0, xo


def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank on the CPU cores next to its GPU (the page-locked staging buffers of the end-to-end
    leg are then allocated on that NUMA node: with 8 ranks the host side of the uploads is the bottleneck)."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores local to {bus}"
    except Exception as e:   # noqa: BLE001 - affinity is an optimisation only
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


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
def leapfrog_bytes_per_chain_step(qn_iters, newton=False, steps_per_launch=1):
    """Algorithmic HBM bytes of one constrained leapfrog step of one chain: doubles
    per SDE time step summed over the phases (X=V=2: a K_t record is 4 doubles, v_t / p_t / x_t records 2)."""
    # momentum projections (pass 1: J p, pass 2: p - J^T lambda), h1 kick and h2_flow fused in:
    project = (4 + 2 + 2 + 2 + 2) + (4 + 2 + 2 + 2 + 2)   # A(dt/2) at the old point + flow -> qw, pw
    project += (4 + 2) + (4 + 2 + 2 + 2 + 2)              # tangent projection at the new point + back flow
    # A(dt/2) at the new point: inside a multi-step launch it is fused with the opening kick of the next step (one
    # projection with a full kick), so it is paid once per launch
    project += ((4 + 2 + 2 + 2 + 2) + (4 + 2 + 2)) / max(steps_per_launch, 1)
    # quasi-Newton: one sweep per iteration (K_t + work position); Newton: forward sweep that also stores the
    # trajectory (4 + 2 + 2) and a backward sweep re-linearising at the iterate (trajectory, work position, K_t)
    qn = (16 if newton else 6) * qn_iters
    qn_final = (4 + 2 + 2 + 2 + 2) + (4 + 2 + 2)           # forward: write q, p; reverse: compare with q_prev
    point = 4 + 8 + 12 + 10                                # trajectory, compressed Jacobian, tangent, adjoint sweeps
    return 8 * T * S * (project + qn + qn_final + point)


def block_sizes(T_, R_, part):
    """Observations per block of partition `part` (mici_extensions.py:321-351)."""
    init = R_ if part == 0 else R_ // 2
    nfull, nrem = divmod(T_ - init, R_)
    nmid = nfull - 1 if nrem == 0 else nfull
    return [init] + [R_] * nmid + [R_ if nrem == 0 else nrem]


def work_model(iters_per_step, T_=T, S_=S, R_=R, X=2, V=2, U=4, Z=4, V0=2):
    """SURVEY.md 8(d) "work model v1": algorithmic FP64 flops (FMA = 2) and the unavoidable HBM bytes of ONE constrained
    leapfrog step of one chain, with the MEASURED number of projection iterations (forward + reverse), averaged over
    the two partitions.  Constants of the model: F_f = 31 (FHN step), F_J = 15, F_Z = 20, F_2 = 100 (second-order
    row-step), adj = 2X^2 + 2XV + 2XZ.  With 5 + 5 iterations and partition 0 this gives the survey's 4,006,468 flops."""
    F_f, F_J, F_Z, F_2 = 31, 15, 20, 100
    adj = 2 * X * X + 2 * X * V + 2 * X * Z
    N = T_ * S_
    dim_q = U + V0 + N * V
    out = []
    for part in (0, 1):
        sizes = block_sizes(T_, R_, part)
        rowsteps = nnzJ = gram = fact = n_c = 0
        for b, n in enumerate(sizes):
            last = b == len(sizes) - 1
            r_b = n if last else (n - 1) + X
            m_b = n * S_ * V + (V0 if b == 0 else 0)
            ks = list(range(1, n + 1)) if last else list(range(1, n)) + [n] * X
            rowsteps += sum(ks) * S_
            nnzJ += r_b * (m_b + U)
            gram += r_b * (r_b + 1) * m_b
            fact += r_b ** 3 / 3 + 2 * r_b ** 2 * U + 2 * r_b * U ** 2
            n_c += r_b
        fact += U ** 3 / 3
        solve = 4 * n_c * (6 + U)
        constr = N * F_f + n_c
        jac = N * (F_f + F_J + F_Z) + rowsteps * adj
        chol_gram = gram + fact
        nsc = 4 * nnzJ + solve
        qn_iter = constr + solve + 2 * nnzJ
        grad_logdet = jac + chol_gram + rowsteps * F_2 + 2 * gram
        out.append(grad_logdet + iters_per_step * qn_iter + 3 * nsc + 20 * dim_q)
    return {"flops": 0.5 * (out[0] + out[1]), "flops_partition0": out[0], "bytes_min": 8 * (4 * dim_q + T_ * X)}


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

    from manifold_mcmc_for_diffusions_b200 import diagnostics, example_models, parallel

    # every rank (also a single one) runs next to its GPU: the page-locked buffers of the end-to-end leg are then
    # allocated on that NUMA node; the original affinity comes back before the CPU baseline uses all the cores
    orig_affinity = os.sched_getaffinity(0)
    affinity = bind_to_gpu_numa_node(local_rank)
    strong = args.scaling == "strong"
    if strong:
        # BASELINE.json config 5: a fixed population sharded over the ranks (contiguous ranges, Philox streams and
        # initial states keyed by the GLOBAL chain index)
        lo, hi = parallel.shard_range(args.total_chains, rank, world)
        n = hi - lo
        y, u, v0, xo = init_inputs_global(lo, hi)
        seed, chain0 = SEED, lo
    else:
        n = args.chains
        y, u, v0, xo = init_inputs(n, rank)
        seed, chain0 = SEED + rank, rank * n
    bc = BatchedChains("fhn", OBS_INTERVAL, S, R, y, 4, n, device=local_rank)
    bc.set_chain_offset(chain0)   # disjoint Philox streams per rank
    bc.opts.solver = 1 if args.solver == "newton" else 0
    if args.regroup:
        bc.set_chain_regrouping(True)   # per-chain results unchanged; chains with similar iteration counts share tiles
    bc.init_linear_interpolation(u, v0, xo, 0)
    del u, v0, xo
    L = args.traj_len
    traces = []

    def trace():
        # the reference scripts' traced scalars (fhn_model_noiseless_obs_chmc_experiment.py:102-117): a D2H read of
        # [u | v_0] of every chain after every transition
        uu, vv = bc.get_head()
        z = example_models.fhn.generate_z(uu)
        traces.append(np.concatenate((z, example_models.fhn.generate_x_0(z, vv)), -1))

    # untimed burn-in towards the typical set (the linear-interpolation states are far in the tails) ...
    n_diag = min(args.diag_transitions, args.burnin) if not args.regroup else 0
    for it in range(args.burnin - n_diag):
        bc.hmc_transition(args.burnin_dt, L, seed, it)
    bc.synchronize()
    # ... whose last transitions run at the bench step size with traces: the draws R-hat / ESS are computed on
    t_diag = time.perf_counter()
    for it in range(args.burnin - n_diag, args.burnin):
        bc.hmc_transition(args.dt, L, seed, it)
        trace()
    bc.synchronize()
    diag_wall = time.perf_counter() - t_diag

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        bc.synchronize()

    it_counter = [args.burnin]

    def do_steps(k, traced=False):
        done = 0
        while done < k:
            m = min(L, k - done)
            bc.transition_begin(seed, it_counter[0])
            bc.transition_steps(args.dt, m)          # m leapfrog steps, one fused launch
            bc.transition_end(seed, it_counter[0], True)
            if traced and m == L and n_diag:
                trace()
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
    do_steps(args.steps, traced=True)       # trace read-backs stay inside the timed region
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

    # page-locked result buffers: the step's RESULT (new q, p of every chain) comes back to the host every step
    outs = [[torch.empty(a.shape, dtype=torch.float64).pin_memory().numpy() for a in pin[:2]] for pin in pins]

    def issue(i):
        qn_, pn_, xn_ = pins[i]
        parts[i].set_state(qn_, xn_, part, p=pn_, blocking=False)   # H2D of this step's inputs (async, own stream)
        parts[i].leapfrog_step(args.dt)
        parts[i].get_state_into(outs[i][0], outs[i][1], None, blocking=False)   # D2H of the new q, p (async)

    def collect(i):
        inf = parts[i].step_info()                # D2H of status / iterations / reverse distance; synchronises
        return int((inf["status"] == 0).sum())

    def e2e_loop(k):
        # software pipeline over the parts: as soon as a part's results are back its next upload + step are queued,
        # so its copies run while the other part computes
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
    assert all(np.isfinite(o[0][0, :8]).all() for o in outs)
    h2d = int(sum(a.nbytes for pin in pins for a in pin))
    d2h = int(sum(a.nbytes for o in outs for a in o) + n * (4 + 4 + 4 + 8))
    if n_parts > 1:
        for b in parts:
            b.close()

    # ---- cross-chain diagnostics: ONE NCCL all-gather of the traced scalars, then rank-normalised split-R-hat and
    # bulk ESS over the whole population on rank 0 (scripts/utils.py:351-381) ----
    diag = None
    if traces:
        tr_local = np.stack(traces, 1)                                   # [chains, draws, 6]
        t_g = time.perf_counter()
        tr_all = parallel.allgather_chains(tr_local.reshape(n, -1))
        if world > 1:
            torch.cuda.synchronize()
        gather_s = time.perf_counter() - t_g
        dw = torch.tensor([diag_wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dw, op=dist.ReduceOp.MAX)
        if rank == 0:
            tr_all = tr_all.reshape(tr_all.shape[0], len(traces), 6)
            t_d = time.perf_counter()
            names = ["sigma", "epsilon", "gamma", "beta", "x_0[0]", "x_0[1]"]
            summ = diagnostics.summary({nm: tr_all[:, :, i] for i, nm in enumerate(names)})
            ess_min = min(v["ess_bulk"] for v in summ.values())
            # ESS/s: effective samples of ALL traced draws over the wall time of the traced transitions, scaled from
            # the separately clocked diagnostic transitions (same step size, trajectory length, traces read back)
            per_tr = dw.item() / max(n_diag, 1)
            diag = {
                "chains": int(tr_all.shape[0]), "draws_per_chain": len(traces),
                "rhat_max": max(v["r_hat"] for v in summ.values()), "ess_bulk_min": ess_min,
                "ess_per_s": ess_min / (per_tr * len(traces) + gather_s),
                "wall_s_per_traced_transition": per_tr,
                "allgather_ms": gather_s * 1e3, "allgather_backend": dist.get_backend() if world > 1 else "none (1 rank)",
                "diagnostics_host_s": time.perf_counter() - t_d,
                "summary": {k: {kk: round(vv, 5) for kk, vv in v.items()} for k, v in summ.items()},
            }

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
        try:
            fp64 = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_b200.json")))
            fp64_peak = float(fp64["fp64_fma_peak_tflops"])
            fp64_src = "measured FMA microbenchmark (profiles/fp64_peak_b200.json)"
        except Exception:
            fp64_peak, fp64_src = 37.2, "nominal 148 SM x 64 DFMA/clk x 1.965 GHz"
        try:
            traffic_rec = json.load(open(os.path.join(ROOT, "profiles", "r2_leapfrog_traffic.json")))
        except Exception:
            traffic_rec = None
        # Roofline of the dominant kernel (k_leapfrog: every phase of the step fused in one launch), three views:
        #  (1) SURVEY.md 8(d) work model: algorithmic FP64 flops with the MEASURED iteration counts vs the measured FP64
        #      peak, and the unavoidable state traffic (read q, p, x_obs, write q, p) vs the measured HBM bandwidth;
        #  (2) this implementation's streaming model (the compressed Jacobian is stored and re-read);
        #  (3) DRAM bytes actually moved (ncu dram__bytes of the committed capture, per successful chain-step).
        # `bound` follows from (1): arithmetic intensity = flops / unavoidable bytes against the ridge point.
        iters_per_step = qn_iters / max(ok, 1)          # forward + reverse iterations per successful step
        wm = work_model(iters_per_step)
        lf_s = ms_lf * 1e-3
        lf_launch_s = lf_s / max(n_lf, 1)
        steps_per_launch = ok / max(n_lf, 1)
        tflops = wm["flops"] * ok / lf_s / 1e12 if lf_s > 0 else 0.0
        floor_gbs = wm["bytes_min"] * ok / lf_s / 1e9 if lf_s > 0 else 0.0
        alg_bytes_step = leapfrog_bytes_per_chain_step(iters_per_step, newton=args.solver == "newton",
                                                       steps_per_launch=args.traj_len)
        impl_gbs = alg_bytes_step * ok / lf_s / 1e9 if lf_s > 0 else 0.0
        intensity = wm["flops"] / wm["bytes_min"]
        ridge = fp64_peak * 1e12 / (hbm_peak * 1e9)
        bound = "fp64" if intensity > ridge else "hbm"
        dram_step = traffic_rec["dram_bytes_per_chain_step"] if traffic_rec else None
        roofline = {
            "bound": bound,
            "kernel": "k_leapfrog (fused constrained leapfrog step: projections, quasi-Newton solves, linearise + grad log-det)",
            "achieved": tflops if bound == "fp64" else floor_gbs,
            "peak": fp64_peak if bound == "fp64" else hbm_peak,
            "unit": "TFLOP/s" if bound == "fp64" else "GB/s",
            "frac": (tflops / fp64_peak) if bound == "fp64" else floor_gbs / hbm_peak,
            "traffic": (dram_step * steps_per_launch) if dram_step else None,
            "traffic_source": traffic_rec["source"] if traffic_rec else None,
            "peak_source": fp64_src if bound == "fp64" else peak_src,
            "bound_reason": "SURVEY 8(d) work model: %.1f flop per unavoidable byte vs ridge %.1f flop/B" % (intensity, ridge),
            "flops_per_chain_step": wm["flops"],
            "alg_bytes_floor_per_chain_step": wm["bytes_min"],
            "alg_bytes_floor_per_launch": wm["bytes_min"] * steps_per_launch,
            "fp64_frac": tflops / fp64_peak,
            "hbm": {
                "peak_gbs": hbm_peak, "peak_source": peak_src,
                "floor_gbs": floor_gbs, "floor_frac": floor_gbs / hbm_peak,
                "impl_model_bytes_per_chain_step": alg_bytes_step, "impl_model_gbs": impl_gbs,
                "impl_model_frac": impl_gbs / hbm_peak,
                "dram_bytes_per_chain_step_ncu": dram_step,
                "dram_gbs": (dram_step * ok / lf_s / 1e9) if dram_step and lf_s > 0 else None,
                "dram_frac": (dram_step * ok / lf_s / 1e9 / hbm_peak) if dram_step and lf_s > 0 else None,
                "traffic_over_floor": (dram_step / wm["bytes_min"]) if dram_step else None,
            },
            "kernel_ms_per_launch": lf_launch_s * 1e3,
            "chain_steps_per_launch": steps_per_launch,
            "launches": n_lf,
            "share_of_timed_region": ms_lf / ms,
        }
        out = {
            "metric": METRIC,
            "value": ok_tot / (ms_max * 1e-3),
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": (WORKLOAD if args.solver == "quasi-newton" else WORKLOAD.replace("quasi-Newton", "Newton"))
                + ("; BASELINE config 5: %d chains sharded over the GPUs" % args.total_chains if strong else ""),
                "total_chains": args.total_chains if strong else n * world,
                "chains_per_gpu": n,
                "cpu_affinity": affinity,
                "chains_per_cta_tile": bc.chains_per_tile(),
                "chain_regrouping": bool(args.regroup),
                "step_size": args.dt,
                "traj_len": L,
                "burnin_transitions": args.burnin,
                "l2": "per-step working set %.1f GB per GPU >> 126 MB L2 (inputs larger than L2, no flush needed)"
                % (n * 450e3 / 1e9),
                "step_success_frac": ok_tot / (n * world * args.steps),
                "accept_stat_last": float(st["accept_stat"].mean()),
                "qn_iters_per_step": iters_per_step,
            },
            "roofline": roofline,
            "e2e": {
                "value": ok_e2e_tot / (e2e_ms_max * 1e-3),
                "unit": UNIT,
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "result_read_back": "new q and p of every chain (page-locked host arrays) + status words, every step",
                "pipeline": "%d BatchedChains objects of %d chains, copies of one overlapped with compute of the other"
                % (n_parts, n // n_parts) if n_parts > 1 else "none",
            },
            "gpu_launches": int(launches_tot),
            "clocks": sampler.summary(),
        }
        if diag is not None:
            out["diagnostics"] = diag
        os.sched_setaffinity(0, orig_affinity)
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
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "total_chains": args.total_chains if args.scaling == "strong" else None,
                   "step_size": args.dt, "traj_len": args.traj_len, "burnin_transitions": args.burnin,
                   "note": "host cores only: every step is a bounded sample of the workload (cpu_baseline.sample)"},
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
    ap.add_argument("--chains", type=int, default=16384, help="chains per GPU (with --scaling weak)")
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
    ap.add_argument("--e2e-pipeline", type=int, default=4,
                    help="number of BatchedChains objects the end-to-end loop alternates between (1 = no overlap)")
    ap.add_argument("--scaling", choices=("weak", "strong"), default="strong",
                    help="strong (default): BASELINE.json config 5, --total-chains sharded over the ranks; weak: --chains "
                         "per GPU")
    ap.add_argument("--total-chains", type=int, default=65536)
    ap.add_argument("--diag-transitions", type=int, default=24,
                    help="last burn-in transitions run at --dt with traces (R-hat / ESS are computed on them and on the "
                         "timed transitions)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
