import os, sys, json, numpy as np
sys.path.insert(0, os.getcwd())
from manifold_mcmc_for_diffusions_b200 import BatchedChains
n=4096; y=np.load("tests/golden/fhn_yseq_T100.npy"); T,S,R=100,25,5
bc=BatchedChains("fhn",0.2,S,R,y,4,n); rng=np.random.default_rng([20200710,0])
u=rng.standard_normal((n,4)); v0=rng.standard_normal((n,2))
xo=np.concatenate((np.broadcast_to(y,(n,T,1)),0.5*rng.standard_normal((n,T,1))),-1)
bc.init_linear_interpolation(u,v0,xo,0)
for it in range(20): bc.hmc_transition(0.05,8,1,it)
bc.transition_begin(1,1000)
res=[]
for s in range(4):
    bc.transition_steps(0.1,1); info=bc.step_info()
    for key in ("iters_fwd","iters_rev"):
        it=info[key].reshape(-1,8)
        res.append((it.mean(), it.max(1).mean(), np.mean([len(np.unique(r)) for r in it])))
res=np.array(res); print(json.dumps({"mean_iters":res[:,0].mean(),"mean_tile_max":res[:,1].mean(),"mean_distinct_convergence_iterations_per_tile":res[:,2].mean()}))
