import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.helpers import make_fhn_problem, make_batched
from oracle import torch_oracle as O
T, S, R = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (10, 5, 5)
part = int(sys.argv[4]) if len(sys.argv) > 4 else 0
prob = make_fhn_problem(T, S, R, n_chains=3, nd=200)
rng = np.random.default_rng(3); p_raw = rng.standard_normal(prob["q"].shape)
sysm = prob["system"]
bc = make_batched(prob)
bc.set_state(prob["q"], prob["xobs"], part, p=p_raw); bc.linearize(True); bc.project_momentum()
rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
ors = []
for i in range(3):
    pt = sysm.point(prob["q"][i], prob["xobs"][i], part)
    p = sysm.project_onto_cotangent_space(torch.tensor(p_raw[i]), pt)
    ors.append([torch.tensor(prob["q"][i]), p, pt])
for s in range(3):
    bc.leapfrog_step(0.05)
    q, p, _ = bc.get_state(); g = bc.grad_log_det_sqrt_gram(); info = bc.step_info()
    for i in range(3):
        qo, po, pto, inf = O.leapfrog_step(sysm, ors[i][0], ors[i][1], prob["xobs"][i], part, 0.05, pt=ors[i][2])
        ors[i] = [qo, po, pto]
        print("step", s, "chain", i, "q rel %.2e p rel %.2e grad rel %.2e" % (rel(q[i], qo.numpy()), rel(p[i], po.numpy()), rel(g[i], pto["grad_ld"].numpy())),
              "iters", info["iters_fwd"][i], inf["n_fwd"], info["iters_rev"][i], inf["n_back"], "status", info["status"][i])
