import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import make_batched, make_fhn_problem
prob = make_fhn_problem(10, 5, 5, n_chains=4, nd=200)
rng = np.random.default_rng(0)
p_raw = rng.standard_normal(prob["q"].shape)
bc = make_batched(prob)
bc.set_state(prob["q"], prob["xobs"], 0, p=p_raw)
bc.linearize(True)
bc.project_momentum()
print("before step", flush=True)
bc.leapfrog_step(0.05)
info = bc.step_info()
print("after step", info["status"], info["iters_fwd"], flush=True)
