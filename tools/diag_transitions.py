import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains
n = int(os.environ.get("NCH", 256))
y = np.load(os.path.join(ROOT, "tests/golden/fhn_yseq_T100.npy")); T, S, R = 100, 25, 5
bc = BatchedChains("fhn", 0.2, S, R, y, 4, n)
rng = np.random.default_rng([20200710, 0])
u = rng.standard_normal((n, 4)); v0 = rng.standard_normal((n, 2))
xo = np.concatenate((np.broadcast_to(y, (n, T, 1)), 0.5 * rng.standard_normal((n, T, 1))), -1)
bc.init_linear_interpolation(u, v0, xo, 0)
print("init |c|", np.abs(bc.constr()).max())
for it in range(6):
    q0, _, _ = bc.get_state()
    bc.transition_begin(1, it)
    print(it, "part", bc.partition, "begin |c|", np.abs(bc.constr()).max(), "H", bc.hamiltonian()[:2])
    for s in range(int(os.environ.get("L", 2))):
        bc.transition_steps(0.05, 1)
        info = bc.step_info()
        print("   step", s, "fail", (info["status"] != 0).mean(), "iters", info["iters_fwd"].mean(), info["iters_rev"].mean(), "|c|", np.abs(bc.constr()).max())
    bc.transition_end(1, it, True)
    st = bc.transition_stats()
    q1, _, x1 = bc.get_state()
    print("   end acc", st["accepted"].mean(), "|c|", np.abs(bc.constr()).max(), "moved", np.abs(q1 - q0).max(1)[:4])
