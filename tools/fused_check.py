import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.helpers import make_fhn_problem, make_batched
prob = make_fhn_problem(10, 5, 5, n_chains=3, nd=200)
rng = np.random.default_rng(3); p_raw = rng.standard_normal(prob["q"].shape)
res = {}
for fused in ("0", "1"):
    os.environ["MMD_FUSED"] = fused
    bc = make_batched(prob)
    bc.set_state(prob["q"], prob["xobs"], 0, p=p_raw); bc.linearize(True); bc.project_momentum()
    out = []
    for s in range(3):
        bc.leapfrog_step(0.05)
        q, p, _ = bc.get_state(); out.append((q, p, bc.step_info(), bc.hamiltonian(), bc.grad_log_det_sqrt_gram()))
    res[fused] = out
    bc.close()
for s in range(3):
    a, b = res["0"][s], res["1"][s]
    print("step", s, "q diff", np.abs(a[0] - b[0]).max(), "p diff", np.abs(a[1] - b[1]).max(), "h diff", np.abs(a[3] - b[3]).max(), "grad diff", np.abs(a[4] - b[4]).max())
    d = np.abs(a[1] - b[1]); print("  p diff rows (chain 0) top:", np.argsort(-d[0])[:8], -np.sort(-d[0])[:4])
    d = np.abs(a[4] - b[4]); print("  g diff rows (chain 0) top:", np.argsort(-d[0])[:8], -np.sort(-d[0])[:4])
