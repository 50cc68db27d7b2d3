"""Times gpk_kbuild alone at C2 (Matern52 fp64 N=8192 D=8): lower-only and full (mirrored) symmetric builds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpflow_b200 as gpf
from gpflow_b200 import ops
from gpflow_b200.kernels import compile_kernel

N, D = 8192, 8
rng = np.random.default_rng(1)
X = torch.as_tensor(rng.standard_normal((N, D)), device="cuda")
k = gpf.kernels.Matern52(lengthscales=np.sqrt(8.0))
desc = compile_kernel(k, D)
K = ops.empty((N, N), like=X)
res = []
for lower in (True, False):
    kw = dict(uplo=gpf._lib.GPK_LOWER) if lower else {}
    for _ in range(3):
        ops.kbuild(desc, X, None, out=K, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.kbuild(desc, X, None, out=K, **kw)
    b.record()
    torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 20 * 1e3)
print("variant", sys.argv[1] if len(sys.argv) > 1 else "-", "lower %.1f us  full %.1f us" % tuple(res))
