import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpflow_b200 as gpf
from oracle import gp_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = O.make_data(2, N, 8, 1)
m = gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.Matern52(lengthscales=np.sqrt(8.0)), noise_variance=0.1)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    v = float(m.log_marginal_likelihood())
print("lml", v)
