import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.test_gpu_sir import make_sir_problem, make_bc
prob = make_sir_problem(6, 4, 6, 3)
bc = make_bc(prob)
bc.set_state(prob["q"], prob["xobs"], 0)
bc.linearize(True)
print("ld", bc.log_det_sqrt_gram())
sysm = prob["system"]
pt = sysm.point(prob["q"][0], prob["xobs"][0], 0)
A = bc.get_factor("A")[0, 0].reshape(16, 5)[:6]
print("A gpu\n", A)
print("A oracle\n", pt["jac"][0][0].numpy())
L = bc.get_factor("L")[0, 0]
print("L packed (first 21)", L[:21])
Lo = pt["chol"][1][0].numpy(); print("chol D oracle\n", Lo)
xe = bc.get_factor("xend")[0, 0].reshape(-1, 3)[:6]; print("xend", xe, "\nxobs", prob["xobs"][0])
K = bc.get_factor("K")[0, 0].reshape(-1, 4, 3, 3)[:6]
print("K[0,0]", K[0, 0], "\nK[5,3]", K[5, 3])
Jv = pt["jac"][1][0].numpy(); print("Jv oracle row0 first cols", Jv[0, :13], Jv.shape)
