"""SURVEY.md 8(f) rank 4 breadth: kernels that are not functions of a Gram term (Cosine, Periodic, ArcCosine, Coregion,
ChangePoints; csrc/kaux.cu), their combinations with fused kernels, GPR on top of them (unfused LML / predict), and the
heteroskedastic Gaussian likelihood.  Device results against the oracle restatements of
gpflow/kernels/stationaries.py:316-332, periodic.py:28-111, misc.py:27-296, changepoints.py:26-193,
likelihoods/scalar_continuous.py:52-148."""
import numpy as np
import pytest

import gpflow_b200 as gpf
from gpflow_b200 import ops
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
K, rng = gpf.kernels, np.random.default_rng(11)
X = rng.standard_normal((150, 3))
X2 = rng.standard_normal((70, 3))
X1 = np.sort(rng.uniform(0, 1, (120, 1)), axis=0)
X1b = rng.uniform(0, 1, (40, 1))

CASES = {
    "cosine": (lambda: K.Cosine(1.3, [0.7, 1.9, 1.1]), lambda: O.Cosine(1.3, [0.7, 1.9, 1.1])),
    "cosine_dims": (lambda: K.Cosine(0.8, 1.4, active_dims=[2, 0]), lambda: O.Cosine(0.8, 1.4, active_dims=[2, 0])),
    "periodic_m32": (lambda: K.Periodic(K.Matern32(1.1, [0.7, 1.2, 0.9]), [1.5, 2.5, 0.8]),
                     lambda: O.Periodic(O.Matern32(1.1, [0.7, 1.2, 0.9]), [1.5, 2.5, 0.8])),
    "periodic_rbf": (lambda: K.Periodic(K.SquaredExponential(1.1, 0.9), 2.0),
                     lambda: O.Periodic(O.SquaredExponential(1.1, 0.9), 2.0)),
    "periodic_m12_dims": (lambda: K.Periodic(K.Matern12(0.6, 0.5, active_dims=[1]), 0.7),
                          lambda: O.Periodic(O.Matern12(0.6, 0.5, active_dims=[1]), 0.7)),
    "arccos0": (lambda: K.ArcCosine(0, 1.2, 1.0, 0.3), lambda: O.ArcCosine(0, 1.2, 1.0, 0.3)),
    "arccos1_ard": (lambda: K.ArcCosine(1, 1.2, [0.5, 2.0, 1.0], 0.3), lambda: O.ArcCosine(1, 1.2, [0.5, 2.0, 1.0], 0.3)),
    "arccos2": (lambda: K.ArcCosine(2, 0.7, 0.4, 1.5), lambda: O.ArcCosine(2, 0.7, 0.4, 1.5)),
    "cosine_plus_rbf_times_m52": (lambda: K.Cosine(1.0, 2.0) + K.SquaredExponential(0.5, 1.5) * K.Matern52(1.0, 2.0),
                                  lambda: O.Cosine(1.0, 2.0) + O.SquaredExponential(0.5, 1.5) * O.Matern52(1.0, 2.0)),
    "periodic_times_linear": (lambda: K.Periodic(K.SquaredExponential(), 1.3) * K.Linear(0.7),
                              lambda: O.Periodic(O.SquaredExponential(), 1.3) * O.Linear(0.7)),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_materialised_kernels_match_oracle(cuda_device, name, dtype):
    kp, ko = CASES[name][0](), CASES[name][1]()
    tol = dict(rtol=1e-11, atol=1e-11) if dtype == np.float64 else dict(rtol=2e-4, atol=2e-5)
    if name.startswith("arccos") and dtype == np.float64:
        # theta = acos(c) has derivative 1 / sqrt(1 - c^2): nearly parallel inputs turn 1e-16 in c into 1e-8 in theta
        # (the reference's own 1e-15 jitter, misc.py:185, exists for this)
        tol = dict(rtol=1e-7, atol=2e-8)
    with gpf.config.as_context(gpf.config.Config(float=dtype)):
        Kxx, Kx2, kd = kp(X), kp(X, X2), kp(X, full_cov=False)
    np.testing.assert_allclose(Kxx.cpu().numpy(), ko(X), **tol)
    np.testing.assert_allclose(Kx2.cpu().numpy(), ko(X, X2), **tol)
    np.testing.assert_allclose(kd.cpu().numpy(), ko(X, full_cov=False), **tol)


def test_coregion_and_changepoints(cuda_device):
    W = rng.standard_normal((3, 2))
    kp = K.Coregion(3, 2, active_dims=[1])
    kp.W.assign(W)
    kp.kappa.assign([0.5, 1.5, 0.7])
    ko = O.Coregion(3, 2, W=W, kappa=[0.5, 1.5, 0.7], active_dims=[1])
    Xc = np.concatenate([rng.standard_normal((60, 1)), rng.integers(0, 3, (60, 1)).astype(float)], axis=1)
    np.testing.assert_allclose(kp(Xc).cpu().numpy(), ko(Xc), rtol=1e-12)
    np.testing.assert_allclose(kp(Xc, full_cov=False).cpu().numpy(), ko(Xc, full_cov=False), rtol=1e-12)
    prod_p = K.SquaredExponential(active_dims=[0]) * kp          # the classic coregionalised regression kernel
    prod_o = O.SquaredExponential(active_dims=[0]) * ko
    np.testing.assert_allclose(prod_p(Xc, Xc[:20]).cpu().numpy(), prod_o(Xc, Xc[:20]), rtol=1e-11)
    with pytest.raises(ValueError):
        K.Coregion(3, 2)(X)                                      # misc.py:262: a 1-D input space
    cp_p = K.ChangePoints([K.SquaredExponential(), K.Matern12(2.0, 0.3), K.Cosine(0.5, 0.2)], [0.3, 0.7], [5.0, 9.0])
    cp_o = O.ChangePoints([O.SquaredExponential(), O.Matern12(2.0, 0.3), O.Cosine(0.5, 0.2)], [0.3, 0.7], [5.0, 9.0])
    np.testing.assert_allclose(cp_p(X1).cpu().numpy(), cp_o(X1), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cp_p(X1, X1b).cpu().numpy(), cp_o(X1, X1b), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cp_p(X1, full_cov=False).cpu().numpy(), cp_o(X1, full_cov=False), rtol=1e-10)
    with pytest.raises(ValueError):
        K.ChangePoints([K.SquaredExponential()], [0.3, 0.5])     # changepoints.py:62-67
    with pytest.raises(ValueError):
        cp_p(X)                                                  # 1-D input space
    with pytest.raises(TypeError):
        K.Periodic(K.Linear())                                   # periodic.py:66-67
    with pytest.raises(ValueError):
        K.ArcCosine(order=3)


def test_gpr_with_materialised_kernel_and_heteroskedastic_noise(cuda_device):
    """GPR LML / predict_f through the unfused composition for a non-fusable kernel, and Gaussian(variance=Function),
    Gaussian(scale=Function) with the reference's lower-bound clip."""
    Y1 = np.sin(6 * X1) + 0.1 * rng.standard_normal(X1.shape)
    kp = K.ChangePoints([K.SquaredExponential(1.0, 0.2), K.Matern32(0.7, 0.1)], [0.5], 20.0)
    ko = O.ChangePoints([O.SquaredExponential(1.0, 0.2), O.Matern32(0.7, 0.1)], [0.5], 20.0)
    m = gpf.models.GPR((X1, Y1), kp, noise_variance=0.05)
    np.testing.assert_allclose(float(m.log_marginal_likelihood()), O.gpr_log_marginal_likelihood(X1, Y1, ko, 0.05), rtol=1e-9)
    mean, var = m.predict_f(X1b)
    mo, vo = O.gpr_predict_f(X1, Y1, ko, 0.05, X1b)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-7, atol=1e-9)
    # heteroskedastic: variance = max(A x + b, 1e-6) and scale = max(..., 1e-3) ** 2
    d = O.make_data(2, 500, 3, 1, n_new=30)
    lin = gpf.mean_functions.Linear(A=np.array([[0.05], [0.0], [-0.02]]), b=np.array([0.08]))
    v_ref = np.maximum(d["X"] @ np.array([[0.05], [0.0], [-0.02]]) + 0.08, 1e-6)
    for kw, vec in (({"variance": lin}, v_ref[:, 0]), ({"scale": lin}, np.maximum(d["X"] @ np.array([[0.05], [0.0], [-0.02]]) + 0.08, 1e-3)[:, 0] ** 2)):
        lik = gpf.likelihoods.Gaussian(**kw)
        assert lik.heteroskedastic
        m = gpf.models.GPR((d["X"], d["Y"]), K.Matern52(lengthscales=2.0), likelihood=lik)
        ref = O.gpr_log_marginal_likelihood(d["X"], d["Y"], O.Matern52(lengthscales=2.0), vec)
        np.testing.assert_allclose(float(m.log_marginal_likelihood()), ref, rtol=1e-9)
        mean, var = m.predict_f(d["Xnew"])
        mo, vo = O.gpr_predict_f(d["X"], d["Y"], O.Matern52(lengthscales=2.0), vec, d["Xnew"])
        np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-7, atol=1e-9)
    # predict_y / predict_log_density add the per-point variance at Xnew (scalar_continuous.py:127-136)
    lik = gpf.likelihoods.Gaussian(variance=lin)
    m = gpf.models.GPR((d["X"], d["Y"]), K.Matern52(lengthscales=2.0), likelihood=lik)
    ym, yv = m.predict_y(d["Xnew"])
    vn = np.maximum(d["Xnew"] @ np.array([[0.05], [0.0], [-0.02]]) + 0.08, 1e-6)
    np.testing.assert_allclose(yv.cpu().numpy(), var.cpu().numpy() * 0 + (m.predict_f(d["Xnew"])[1].cpu().numpy() + vn), rtol=1e-10)
    with pytest.raises(NotImplementedError):
        gpf.models.SGPR((d["X"], d["Y"]), K.Matern52(), d["X"][:10], likelihood=lik).elbo()
