import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from manifold_mcmc_for_diffusions_b200 import BatchedChains
gold = np.load(os.path.join(ROOT, "tests/golden/fhn_T100_S25_R5_golden.npz"), allow_pickle=True)
part = int(os.environ.get("PART", 1)); rep = int(os.environ.get("REP", 8))
idx = [c for c in range(gold["q0"].shape[0]) if c % 2 == part]
q0 = np.tile(gold["q0"][idx], (rep, 1)); xo = np.tile(gold["xobs"][idx], (rep, 1, 1)); p = np.tile(gold["p_raw"][idx], (rep, 1))
n = q0.shape[0]; m = len(idx)
bc = BatchedChains("fhn", 0.2, int(gold["S"]), int(gold["R"]), gold["y"], 4, n)
via = os.environ.get("VIA", "set")
if via == "set":
    bc.set_state(q0, xo, part, p=p)
else:
    bc.set_state(q0, xo, 1 - part, p=p); bc.switch_partition(); bc.set_momentum(p)
def dev(a): a = a.reshape(rep, m, -1); return np.abs(a - a[0]).max()
q, pp, x = bc.get_state(); print("roundtrip q", np.abs(q - q0).max(), "x", np.abs(x - xo).max())
print("constr dev", dev(bc.constr()))
bc.linearize(True); print("ld dev", dev(bc.log_det_sqrt_gram()), "grad dev", dev(bc.grad_log_det_sqrt_gram()))
bc.project_momentum(); _, pp, _ = bc.get_state(); print("p dev", dev(pp))
bc.leapfrog_step(float(gold["dt"])); info = bc.step_info(); q, pp, _ = bc.get_state()
print("status", info["status"].reshape(rep, m), "iters", info["iters_fwd"].reshape(rep, m)[:, 0], "q dev", dev(q))
