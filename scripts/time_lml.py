"""Times GPR.log_marginal_likelihood() at the C2 shape (or N given) with CUDA events; prints ms/eval, the value and the
per-class kernel times.  Used for A/B runs of env-var selected kernel variants (one process per setting)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpflow_b200 as gpf
from gpflow_b200 import _lib
from oracle import gp_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tag = sys.argv[3] if len(sys.argv) > 3 else ""
lib = _lib.load()
d = O.make_data(2, N, 8, 1)
m = gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.Matern52(lengthscales=np.sqrt(8.0)), noise_variance=0.1)
for _ in range(3):
    v = m.log_marginal_likelihood()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib.gpk_launch_count_reset()
e0.record()
for _ in range(steps):
    v = m.log_marginal_likelihood()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
launches = lib.gpk_launch_count() / steps
lib.gpk_prof_enable(1)
for _ in range(steps):
    m.log_marginal_likelihood()
msv = (ctypes.c_double * 5)(); cnt = (ctypes.c_int64 * 5)()
lib.gpk_prof_read(msv, cnt, 5)
lib.gpk_prof_enable(0)
cls = {k: round(msv[i] / steps, 3) for i, k in enumerate(["kbuild", "gemm", "leaf", "skinny", "misc"])}
print(f"[{tag}] N={N} ms/eval={ms:.3f} launches={launches:.0f} lml={float(v):.10f} classes={cls}", flush=True)
