"""Generates tests/golden/*.npz|json from the CPU oracle (oracle/gp_oracle.py).

The reference (GPflow on TensorFlow) cannot be imported in this image — TensorFlow is absent — so
these fixtures are ORACLE outputs; the oracle itself is pinned to the reference's known-answer tests
in tests/test_oracle_pins.py.  Inputs follow SURVEY.md 8(d): rng = default_rng(20220523 + c).

    python tests/golden/make_golden.py            # reduced-size fixtures (seconds)
    python tests/golden/make_golden.py --full     # + full-size scalars for configs C2..C5 (minutes)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def kernels_for(c, D):
    s = np.sqrt(D)
    if c == 1:
        return O.SquaredExponential(variance=1.0, lengthscales=s)
    if c == 2:
        return O.Matern52(variance=1.0, lengthscales=s)
    if c == 3:
        return O.SquaredExponential(variance=1.0, lengthscales=s)
    if c == 4:
        return O.SquaredExponential(variance=1.0, lengthscales=s) + O.White(variance=0.1)
    raise ValueError(c)


def c5_kernels(D, P=4):
    s = np.sqrt(D)
    return [(O.SquaredExponential(variance=1.0 + 0.1 * p, lengthscales=s * (1 + 0.05 * p))
             + O.Matern32(variance=1.0, lengthscales=2 * s)) * O.Linear(variance=1.0 / (1 + p)) for p in range(P)]


def small():
    out = {}
    # C1: GPR RBF fp64 N=512 D=2 (full size already)
    d = O.make_data(1, 512, 2, 1, n_new=64)
    k = kernels_for(1, 2)
    out["c1_lml"] = O.gpr_log_marginal_likelihood(d["X"], d["Y"], k, 0.1)
    m, v = O.gpr_predict_f(d["X"], d["Y"], k, 0.1, d["Xnew"])
    out["c1_mean"], out["c1_var"] = m, v
    # C2 reduced: N=1024 D=8
    d = O.make_data(2, 1024, 8, 1, n_new=32)
    k = kernels_for(2, 8)
    out["c2_lml"] = O.gpr_log_marginal_likelihood(d["X"], d["Y"], k, 0.1)
    m, v = O.gpr_predict_f(d["X"], d["Y"], k, 0.1, d["Xnew"])
    out["c2_mean"], out["c2_var"] = m, v
    # C3 reduced: N=5000 M=256 D=16 fp32, jitter 1e-4
    d = O.make_data(3, 5000, 16, 1, M=256, n_new=100, dtype=np.float32)
    k = kernels_for(3, 16)
    out["c3_elbo"] = O.sgpr_elbo(d["X"], d["Y"], k, d["Z"], 0.1, jitter=1e-4)
    d64 = O.make_data(3, 5000, 16, 1, M=256, n_new=100, dtype=np.float64)
    out["c3_elbo_f64"] = O.sgpr_elbo(d64["X"], d64["Y"], k, d64["Z"], 0.1, jitter=1e-4)
    m, v = O.sgpr_predict_f(d64["X"], d64["Y"], k, d64["Z"], 0.1, d64["Xnew"], jitter=1e-4)
    out["c3_mean_f64"], out["c3_var_f64"] = m, v
    # C4 reduced: N=20000 B=512 M=128 P=4 D=16 fp32
    d64 = O.make_data(4, 20000, 16, 4, M=128, dtype=np.float64)
    k = kernels_for(4, 16)
    q_mu, q_sqrt = O.make_q(4, 128, 4)
    Xb, Yb = d64["X"][:512], d64["Y"][:512]
    out["c4_elbo_f64"] = O.svgp_elbo(Xb, Yb, d64["Z"], k, q_mu, q_sqrt, 0.1, whiten=True, num_data=20000, jitter=1e-4)
    out["c4_elbo_nowhite_f64"] = O.svgp_elbo(Xb, Yb, d64["Z"], k, q_mu, q_sqrt, 0.1, whiten=False, num_data=20000,
                                             jitter=1e-4)
    # C5 reduced: N=512 D=32, 4 outputs
    d = O.make_data(5, 512, 32, 4)
    out["c5_lml"] = O.separate_gpr_lml(d["X"], d["Y"], c5_kernels(32), 0.1)
    np.savez(os.path.join(HERE, "golden_small.npz"), **{k_: np.asarray(v_) for k_, v_ in out.items()})
    print({k_: (float(v_) if np.ndim(v_) == 0 else np.shape(v_)) for k_, v_ in out.items()})


def full():
    res = {}
    d = O.make_data(2, 8192, 8, 1)
    res["c2_lml_N8192_D8_f64"] = O.gpr_log_marginal_likelihood(d["X"], d["Y"], kernels_for(2, 8), 0.1)
    print(res)
    d = O.make_data(3, 100000, 16, 1, M=1024)
    res["c3_elbo_N100000_M1024_D16_f64_jitter1e-4"] = O.sgpr_elbo(d["X"], d["Y"], kernels_for(3, 16), d["Z"], 0.1,
                                                                    jitter=1e-4)
    print(res)
    d = O.make_data(4, 1000000, 16, 8, M=2048)
    q_mu, q_sqrt = O.make_q(4, 2048, 8)
    res["c4_elbo_N1e6_B4096_M2048_P8_D16_f64_jitter1e-4_batch0"] = O.svgp_elbo(
        d["X"][:4096], d["Y"][:4096], d["Z"], kernels_for(4, 16), q_mu, q_sqrt, 0.1, whiten=True, num_data=1000000,
        jitter=1e-4)
    print(res)
    d = O.make_data(5, 4096, 32, 4)
    res["c5_lml_N4096_D32_P4_f64"] = O.separate_gpr_lml(d["X"], d["Y"], c5_kernels(32), 0.1)
    print(res)
    json.dump(res, open(os.path.join(HERE, "golden_full.json"), "w"), indent=1)


def c3_predict_full():
    """BASELINE configs[2] "posterior predict" at full size: SGPR posterior mean / variance at Xnew [10000, 16]
    (gpflow/posteriors.py:479-551).  Stored: the first 256 rows and the sums over all 10000 (size-independent checksum)."""
    d = O.make_data(3, 100000, 16, 1, M=1024, n_new=10000)
    m, v = O.sgpr_predict_f(d["X"], d["Y"], kernels_for(3, 16), d["Z"], 0.1, d["Xnew"], jitter=1e-4)
    np.savez(os.path.join(HERE, "golden_c3_predict.npz"), mean=m[:256], var=v[:256], mean_sum=m.sum(), var_sum=v.sum(),
             mean_abs_sum=np.abs(m).sum())
    print("c3 predict", m[:3, 0], v[:3, 0], m.sum(), v.sum())


if __name__ == "__main__":
    if "--c3-predict" in sys.argv:
        c3_predict_full()
        sys.exit(0)
    small()
    if "--full" in sys.argv:
        full()
