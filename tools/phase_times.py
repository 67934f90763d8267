#!/usr/bin/env python
"""Where the fused leapfrog kernel spends its time: per-phase cycle counters of thread 0 of every CTA, from a
build with -DMMD_PHASE_CLOCK (tools/build_variant.sh phase -DMMD_PHASE_CLOCK; MMD_B200_LIB=build_variants/libmmd_phase.so).
Cycles are CTA-resident cycles (several CTAs share an SM), so only the SHARES are meaningful."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains  # noqa: E402

n = int(os.environ.get("NCH", 16384)); burn = int(os.environ.get("BURN", 30)); dt = float(os.environ.get("DT", 0.1))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy"))
T, S, R = 100, 25, 5
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
bc.opts.solver = int(os.environ.get("SOLVER", 0))
rng = np.random.default_rng([20200710, 0])
u = rng.standard_normal((n, 4)); v0 = rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
for it in range(burn):
    bc.hmc_transition(0.05, 8, 1, it)
out = (C.c_ulonglong * 64)()
from manifold_mcmc_for_diffusions_b200._lib import check  # noqa: E402
check(bc._L.mmd_debug_phase_cycles(bc._h, out, 1))
bc.successful_steps(reset=True)
bc.timer_start()
L, ntr = 8, 2
for tr in range(ntr):
    bc.transition_begin(1, 1000 + tr)
    bc.transition_steps(dt, L)
    bc.transition_end(1, 1000 + tr, True)
ms = bc.timer_stop_ms()
check(bc._L.mmd_debug_phase_cycles(bc._h, out, 0))
c = np.array(list(out), dtype=np.float64)
step_cycles = c[0:7].sum()
names = {0: "project: half kick + cotangent projection + h2 flow", 1: "projection solve, forward", 2: "linearise + grad log det",
         3: "project: reverse flow", 4: "projection solve, reverse check", 5: "project: second half kick", 6: "commit"}
res = {"chains": n, "ms_per_step": ms / (L * ntr), "chain_steps_per_s": bc.successful_steps() / (ms * 1e-3),
       "cta_steps": c[7], "solver_iterations_per_cta_step": c[12] / c[7], "check_passes_per_cta_step": c[13] / c[7],
       "exact_norm_passes_per_cta_step": c[14] / c[7],
       "share_of_step": {names[i]: round(c[i] / step_cycles, 4) for i in range(7)},
       "inside_solves_share_of_step": {"sweep (incl. lock-step wait)": round(c[9] / step_cycles, 4),
                                       "block solves + multiplier recursion": round(c[10] / step_cycles, 4),
                                       "convergence checks (bound, exact passes)": round(c[11] / step_cycles, 4),
                                       "deferred iterate write": round(c[15] / step_cycles, 4),
                                       "loop top": round(c[8] / step_cycles, 4)},
       "inside_linearise_share_of_step": {"sweeps 1+2 (trajectory, compressed Jacobian)": round(c[16] / step_cycles, 4),
                                          "per-observation algebra + Cholesky": round(c[17] / step_cycles, 4),
                                          "capacitance reduce": round(c[18] / step_cycles, 4),
                                          "grad: block algebra": round(c[19] / step_cycles, 4),
                                          "grad: per-interval setup": round(c[22] / step_cycles, 4),
                                          "grad: forward tangent sweep": round(c[20] / step_cycles, 4),
                                          "grad: reverse adjoint sweep": round(c[21] / step_cycles, 4),
                                          "grad: tail + reduce": round(c[23] / step_cycles, 4)}}
# absolute CTA-resident cycles per tile and step (meaningful as latencies when one CTA is resident per SM: NCH <= 148 * cpb)
keys = {0: "project_fwd", 1: "solve_fwd", 2: "linearise", 3: "project_rev", 4: "solve_rev", 5: "project_kick2", 6: "commit", 9: "solve_sweeps",
        10: "solve_block_algebra", 11: "solve_checks", 15: "solve_final_pass", 16: "lin_sweeps12", 17: "lin_obs_algebra", 18: "lin_cap_reduce",
        19: "grad_block_algebra", 22: "grad_setup", 20: "grad_fwd_sweep", 21: "grad_rev_sweep", 23: "grad_tail"}
res["cycles_per_cta_step"] = {v: round(c[k] / c[7]) for k, v in keys.items()}
res["solver_iteration_detail_cycles_per_iteration"] = {
    "thread0_working_fraction": round(c[30] / max(c[12], 1), 3),
    "loop top: parameters, coefficients, block start": round(c[24] / max(c[12], 1)),
    "sweep call": round(c[25] / max(c[12], 1)), "  of which waits for the ring (cp.async)": round(c[31] / max(c[12], 1)),
    "  of which the serial recursion": round(c[8] / max(c[12], 1)), "barrier after sweep": round(c[26] / max(c[12], 1)),
    "Woodbury block solve incl. cross-block reduction": round(c[27] / max(c[12], 1)),
    "  first product (factor loads, D^-1 r, (D^-1 A)^T r, capacitance factor loads)": round(c[32] / max(c[12], 1)),
    "  cross-block reduction (barrier + sums)": round(c[33] / max(c[12], 1)),
    "  capacitance solve": round(c[34] / max(c[12], 1)), "  second product": round(c[35] / max(c[12], 1)),
    "multiplier update + alpha recursion": round(c[28] / max(c[12], 1)), "barrier (any check)": round(c[29] / max(c[12], 1))}
res["step_cycles"] = round(step_cycles / c[7])
print(json.dumps(res))
