#!/usr/bin/env python
"""Simulate the canonical FHN observation sequence (fhn_model_noiseless_obs_chmc_experiment.py:84-93:
seed 20200710, z=[0.3,0.1,1.5,0.8], x_0=[-0.5,0.2], obs_interval 0.2, 10,000 fine steps per obs,
T=1000 so every op-timing grid size is a prefix) and store it as a small fixture."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.models import fhn_simulate_y_seq_numpy

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nd = 10000
rng = np.random.default_rng(20200710)
v = rng.standard_normal((T * nd, 2))
y = fhn_simulate_y_seq_numpy(np.array([0.3, 0.1, 1.5, 0.8]), np.array([-0.5, 0.2]), v, 0.2 / nd, nd)
out = os.path.join(ROOT, "tests", "golden", f"fhn_yseq_T{T}.npy")
np.save(out, y)
print("wrote", out, y.shape, y[:3].ravel())
