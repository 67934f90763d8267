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
out = (C.c_ulonglong * 32)()
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
print(json.dumps(res))
