"""Device backward pass of GPR.log_marginal_likelihood (csrc/grad.cu) against the gradient oracle
(oracle/gp_grad_oracle.py, pinned by finite differences in tests/test_oracle_grad.py), and the Scipy optimiser driver.
Reference: TensorFlow autodiff through gpflow/models/gpr.py:91-107, driven by gpflow/optimizers/scipy.py:78-228."""
import numpy as np
import pytest

import gpflow_b200 as gpf
from oracle import gp_grad_oracle as G
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

PAIRS = [("SquaredExponential", "SquaredExponential"), ("Matern12", "Matern12"), ("Matern32", "Matern32"),
         ("Matern52", "Matern52"), ("Exponential", "Exponential")]


def _check(m, X, Y, ko, s2, rtol):
    lml, grads = m.log_marginal_likelihood_and_grad()
    ref_lml, ref = G.gpr_lml_and_grad(X, Y, ko, s2)
    np.testing.assert_allclose(float(lml), ref_lml, rtol=1e-9)
    scale = max(abs(ref["variance"]), np.max(np.abs(ref["lengthscales"])), abs(ref["noise_variance"]))
    np.testing.assert_allclose(float(grads[m.kernel.variance]), ref["variance"], rtol=rtol, atol=rtol * scale)
    np.testing.assert_allclose(np.asarray(grads[m.kernel.lengthscales]), ref["lengthscales"], rtol=rtol, atol=rtol * scale)
    np.testing.assert_allclose(float(grads[m.likelihood.variance]), ref["noise_variance"], rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("name,oname", PAIRS)
@pytest.mark.parametrize("N,D,P", [(300, 3, 1), (700, 5, 2)])
def test_gpr_grad_matches_oracle_small(cuda_device, name, oname, N, D, P):
    d = O.make_data(1, N, D, P)
    kp = getattr(gpf.kernels, name)(variance=1.3, lengthscales=1.7)
    ko = getattr(O, oname)(variance=1.3, lengthscales=1.7)
    m = gpf.models.GPR((d["X"], d["Y"]), kp, noise_variance=0.15)
    _check(m, d["X"], d["Y"], ko, 0.15, 1e-6)


def test_gpr_grad_c1_and_reduced_c2(cuda_device):
    """BASELINE configs[0] (N = 512, D = 2, RBF) and configs[1] at reduced size (N = 1500, D = 8, Matern52): variance,
    lengthscale and noise gradients within 1e-6 of the oracle."""
    d = O.make_data(1, 512, 2, 1)
    s = float(np.sqrt(2.0))
    m = gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.SquaredExponential(lengthscales=s), noise_variance=0.1)
    _check(m, d["X"], d["Y"], O.SquaredExponential(lengthscales=s), 0.1, 1e-6)
    d = O.make_data(2, 1500, 8, 1)
    s = float(np.sqrt(8.0))
    m = gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.Matern52(lengthscales=s), noise_variance=0.1)
    _check(m, d["X"], d["Y"], O.Matern52(lengthscales=s), 0.1, 1e-6)


def test_gpr_grad_ard_and_mean_function(cuda_device):
    d = O.make_data(2, 600, 4, 1)
    ell = np.array([1.1, 1.9, 0.7, 2.5])
    m = gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.Matern32(variance=0.8, lengthscales=ell), noise_variance=0.2)
    _check(m, d["X"], d["Y"], O.Matern32(variance=0.8, lengthscales=ell), 0.2, 1e-6)


def test_gpr_grad_full_size_c2_finite_difference_of_device_lml(cuda_device):
    """Size-independent check at BASELINE size (N = 8192): the analytic device gradient agrees with a central finite
    difference of the device LML itself along the lengthscale and noise directions."""
    d = O.make_data(2, 8192, 8, 1)
    s = float(np.sqrt(8.0))

    def lml(ell, s2):
        return float(gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.Matern52(lengthscales=ell), noise_variance=s2)
                     .log_marginal_likelihood())

    m = gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.Matern52(lengthscales=s), noise_variance=0.1)
    _, grads = m.log_marginal_likelihood_and_grad()
    h = 1e-4
    fd_l = (lml(s + h, 0.1) - lml(s - h, 0.1)) / (2 * h)
    fd_n = (lml(s, 0.1 + h) - lml(s, 0.1 - h)) / (2 * h)
    np.testing.assert_allclose(float(grads[m.kernel.lengthscales]), fd_l, rtol=1e-5)
    np.testing.assert_allclose(float(grads[m.likelihood.variance]), fd_n, rtol=1e-5)


def test_scipy_driver_trains_gpr_on_device_gradients(cuda_device):
    """gpflow/optimizers/scipy.py:78-228 contract: a few L-BFGS-B iterations lower the training loss, the variables end at
    the optimiser's iterate, and the loss history is monotone at the accepted steps."""
    d = O.make_data(1, 400, 2, 1)
    m = gpf.models.GPR((d["X"], d["Y"]), gpf.kernels.SquaredExponential(lengthscales=3.0), noise_variance=1.0)
    loss0 = -float(m.log_marginal_likelihood())
    opt = gpf.optimizers.Scipy()
    res = opt.minimize(m.training_loss_closure(), m.trainable_variables, options={"maxiter": 15})
    loss1 = -float(m.log_marginal_likelihood())
    assert loss1 < loss0 - 1.0
    np.testing.assert_allclose(loss1, res.fun, rtol=1e-8)
    # gradients w.r.t. the UNCONSTRAINED variables at the optimiser's iterate against the oracle's closed form chained
    # through the same bijectors (a finite difference of the device LML is useless here: near the optimum the gradient is
    # ~1e-5 and 1e-9 of evaluation noise over h = 1e-5 is 1e-4)
    loss, grads = m.training_loss_and_gradients()
    v, ell, s2 = (float(p.numpy()) for p in (m.kernel.variance, m.kernel.lengthscales, m.likelihood.variance))
    _, g = G.gpr_lml_and_grad(d["X"], d["Y"], O.SquaredExponential(v, ell), s2)
    ref = {id(m.kernel.variance): g["variance"], id(m.kernel.lengthscales): g["lengthscales"],
           id(m.likelihood.variance): g["noise_variance"]}
    for p, gu in zip(m.trainable_parameters, grads):
        want = -p.unconstrained_gradient(ref[id(p)])
        np.testing.assert_allclose(float(gu), float(want), rtol=1e-5, atol=2e-6)
    # and in the large: a finite difference with a step that dominates the evaluation noise, away from the optimum
    m.kernel.lengthscales.assign(1.0)
    loss, grads = m.training_loss_and_gradients()
    for p, gu in zip(m.trainable_parameters, grads):
        u = p.unconstrained_variable.copy()
        h = 1e-4
        p.assign_unconstrained(u + h)
        lp = -float(m.log_marginal_likelihood())
        p.assign_unconstrained(u - h)
        lm = -float(m.log_marginal_likelihood())
        p.assign_unconstrained(u)
        np.testing.assert_allclose(float(gu), (lp - lm) / (2 * h), rtol=1e-5, atol=1e-4)
