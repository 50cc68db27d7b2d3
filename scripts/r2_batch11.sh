#!/bin/bash
# Round-2 batch 11: compute-sanitizer memcheck on the smoke path and on a factorisation that engages every new kernel
# (fused panel with critical CTAs, plane emission, tcgen05 update from the plane store, gradient pass).
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== sanitizer: smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py --smoke > gpurun_out/b11_san_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/b11_san_smoke.log
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import gpflow_b200 as gpf
from oracle import gp_oracle as O, gp_grad_oracle as G
d = O.make_data(2, 1100, 5, 2)
kp, ko = gpf.kernels.Matern52(lengthscales=2.0), O.Matern52(lengthscales=2.0)
m = gpf.models.GPR((d["X"], d["Y"]), kp, noise_variance=0.1)
v = float(m.log_marginal_likelihood())
ref = O.gpr_log_marginal_likelihood(d["X"], d["Y"], ko, 0.1)
assert abs(v - ref) < 1e-8 * abs(ref), (v, ref)
lml, g = m.log_marginal_likelihood_and_grad()
_, gr = G.gpr_lml_and_grad(d["X"], d["Y"], ko, 0.1)
assert abs(float(g[m.kernel.lengthscales]) - gr["lengthscales"]) < 1e-5 * abs(gr["lengthscales"]) + 1e-6
k2 = gpf.kernels.Cosine(1.0, 2.0) + gpf.kernels.SquaredExponential()
print("ok", v, k2(d["X"][:50]).shape)
PY
echo "== sanitizer: N=1100 GPR value + grad + materialised kernel"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san_case.py > gpurun_out/b11_san_case.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/b11_san_case.log
