"""SURVEY.md 8(f) ranks 2-3 on the device against the oracle: predict_f_samples (model.py:232-288, util.py:179-211) with
injected standard-normal draws, GPRFITC (sgpr.py:380-523), and the multi-output SVGP posteriors
(posteriors.py:844-901: independent latents with shared / separate kernels and inducing variables,
LinearCoregionalization mixing, all four full_cov x full_output_cov forms)."""
import numpy as np
import pytest

import gpflow_b200 as gpf
from gpflow_b200.inducing_variables import SeparateIndependentInducingVariables, SharedIndependentInducingVariables
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
K = gpf.kernels


def test_predict_f_samples_with_injected_draws(cuda_device):
    d = O.make_data(1, 300, 2, 2, n_new=12)
    rng = np.random.default_rng(3)
    m = gpf.models.GPR((d["X"], d["Y"]), K.Matern52(lengthscales=1.5), noise_variance=0.1)
    mo, vo = O.gpr_predict_f(d["X"], d["Y"], O.Matern52(lengthscales=1.5), 0.1, d["Xnew"], full_cov=True)
    eps = rng.standard_normal((2, 12, 5))                      # [P, N, S]
    got = m.predict_f_samples(d["Xnew"], 5, full_cov=True, eps=eps)
    assert tuple(got.shape) == (5, 12, 2)
    np.testing.assert_allclose(got.cpu().numpy(), O.predict_f_samples(mo, vo, True, eps), rtol=1e-7, atol=1e-8)
    mo, vo = O.gpr_predict_f(d["X"], d["Y"], O.Matern52(lengthscales=1.5), 0.1, d["Xnew"], full_cov=False)
    eps = rng.standard_normal((4, 12, 2))                      # [S, N, P]
    got = m.predict_f_samples(d["Xnew"], 4, full_cov=False, eps=eps)
    np.testing.assert_allclose(got.cpu().numpy(), O.predict_f_samples(mo, vo, False, eps), rtol=1e-8, atol=1e-9)
    one = m.predict_f_samples(d["Xnew"], None, full_cov=False, eps=eps[:1])
    assert tuple(one.shape) == (12, 2)
    drawn = m.predict_f_samples(d["Xnew"], 3)                  # library draws: shape and finiteness
    assert tuple(drawn.shape) == (3, 12, 2) and np.isfinite(drawn.cpu().numpy()).all()
    with pytest.raises(NotImplementedError):
        m.predict_f_samples(d["Xnew"], 2, full_cov=True, full_output_cov=True)


@pytest.mark.parametrize("P", [1, 2])
def test_gprfitc_matches_oracle(cuda_device, P):
    d = O.make_data(3, 1200, 4, P, M=60, n_new=25)
    kp, ko = K.SquaredExponential(1.2, 1.7), O.SquaredExponential(1.2, 1.7)
    m = gpf.models.GPRFITC((d["X"], d["Y"]), kp, d["Z"], noise_variance=0.15)
    np.testing.assert_allclose(float(m.fitc_log_marginal_likelihood()), O.gprfitc_lml(d["X"], d["Y"], ko, d["Z"], 0.15),
                               rtol=1e-9)
    np.testing.assert_allclose(float(m.maximum_log_likelihood_objective()),
                               O.gprfitc_lml(d["X"], d["Y"], ko, d["Z"], 0.15), rtol=1e-9)
    for full_cov in (False, True):
        mean, var = m.predict_f(d["Xnew"], full_cov=full_cov)
        mo, vo = O.gprfitc_predict_f(d["X"], d["Y"], ko, d["Z"], 0.15, d["Xnew"], full_cov=full_cov)
        np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-7, atol=1e-9)
    # Z = X: FITC is exact (sgpr.py docstring; tests/gpflow/models/test_sgpr.py)
    sub = slice(0, 200)
    g = gpf.models.GPR((d["X"][sub], d["Y"][sub]), kp, noise_variance=0.15)
    f = gpf.models.GPRFITC((d["X"][sub], d["Y"][sub]), kp, d["X"][sub], noise_variance=0.15)
    np.testing.assert_allclose(float(f.fitc_log_marginal_likelihood()), float(g.log_marginal_likelihood()), rtol=1e-5)


def _mo_setup(L, P=None):
    rng = np.random.default_rng(9)
    d = O.make_data(4, 400, 3, P or L, M=25, n_new=15)
    q_mu = 0.3 * rng.standard_normal((25, L))
    q_sqrt = np.stack([np.tril(0.1 * rng.standard_normal((25, 25))) + 0.7 * np.eye(25) for _ in range(L)])
    return d, q_mu, q_sqrt, rng


@pytest.mark.parametrize("whiten", [True, False])
@pytest.mark.parametrize("full_cov", [False, True])
def test_independent_multioutput_posteriors(cuda_device, whiten, full_cov):
    d, q_mu, q_sqrt, rng = _mo_setup(3)
    kps = [K.SquaredExponential(1.0, 1.2), K.Matern32(0.7, 2.0), K.Matern52(1.3, 0.8)]
    kos = [O.SquaredExponential(1.0, 1.2), O.Matern32(0.7, 2.0), O.Matern52(1.3, 0.8)]
    Zs = [d["Z"], d["Z"] + 0.1, d["Z"] - 0.2]
    cases = {
        "shared_shared": (K.SharedIndependent(kps[0], 3), SharedIndependentInducingVariables(d["Z"]), [d["Z"]], [kos[0]]),
        "separate_shared": (K.SeparateIndependent(kps), SharedIndependentInducingVariables(d["Z"]), [d["Z"]], kos),
        "shared_separate": (K.SharedIndependent(kps[0], 3), SeparateIndependentInducingVariables(Zs), Zs, [kos[0]]),
        "separate_separate": (K.SeparateIndependent(kps), SeparateIndependentInducingVariables(Zs), Zs, kos),
    }
    for name, (kern, iv, zs, ks) in cases.items():
        m = gpf.models.SVGP(kern, gpf.likelihoods.Gaussian(0.1), iv, num_latent_gps=3, q_mu=q_mu, q_sqrt=q_sqrt,
                            whiten=whiten, num_data=400)
        mean, var = m.predict_f(d["Xnew"], full_cov=full_cov)
        mo, vo = O.mo_independent_predict_f(d["Xnew"], zs, ks, q_mu, q_sqrt, whiten=whiten, full_cov=full_cov)
        np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-7, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-7, atol=1e-9, err_msg=name)
        if not full_cov:                                       # full_output_cov: diagonal over outputs (util.py:222-254)
            _, v3 = m.predict_f(d["Xnew"], full_cov=False, full_output_cov=True)
            ref = np.zeros((15, 3, 3))
            ref[:, np.arange(3), np.arange(3)] = vo
            np.testing.assert_allclose(v3.cpu().numpy(), ref, rtol=1e-7, atol=1e-9, err_msg=name)
    # the ELBO of a multi-output model composes prior_kl + predict_f + variational expectations (svgp.py:166-181)
    kern, iv, zs, ks = cases["separate_separate"]
    m = gpf.models.SVGP(kern, gpf.likelihoods.Gaussian(0.1), iv, num_latent_gps=3, q_mu=q_mu, q_sqrt=q_sqrt,
                        whiten=whiten, num_data=400)
    Xb, Yb = d["X"][:64], d["Y"][:64]
    fm, fv = O.mo_independent_predict_f(Xb, zs, ks, q_mu, q_sqrt, whiten=whiten)
    kl = sum(O.gauss_kl(q_mu[:, l:l + 1], q_sqrt[l:l + 1], None if whiten else O.Kuu(zs[l], ks[l], jitter=1e-6))
             for l in range(3))
    ref = float(np.sum(O.gaussian_variational_expectations(fm, fv, Yb, 0.1)) * 400 / 64 - kl)
    np.testing.assert_allclose(float(m.elbo((Xb, Yb))), ref, rtol=1e-8)


@pytest.mark.parametrize("full_cov,full_output_cov", [(False, False), (False, True), (True, False), (True, True)])
def test_linear_coregionalization_posterior(cuda_device, full_cov, full_output_cov):
    d, q_mu, q_sqrt, rng = _mo_setup(2, P=3)
    W = rng.standard_normal((3, 2))
    kps, kos = [K.SquaredExponential(1.0, 1.2), K.Matern32(0.7, 2.0)], [O.SquaredExponential(1.0, 1.2), O.Matern32(0.7, 2.0)]
    kern = K.LinearCoregionalization(kps, W)
    m = gpf.models.SVGP(kern, gpf.likelihoods.Gaussian(0.1), SharedIndependentInducingVariables(d["Z"]), num_latent_gps=2,
                        q_mu=q_mu, q_sqrt=q_sqrt, whiten=True, num_data=400)
    mean, cov = m.predict_f(d["Xnew"], full_cov=full_cov, full_output_cov=full_output_cov)
    gm, gv = O.mo_independent_predict_f(d["Xnew"], [d["Z"]], kos, q_mu, q_sqrt, whiten=True, full_cov=full_cov)
    mo, co = O.mix_latent_gp(W, gm, gv, full_cov, full_output_cov)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(cov.cpu().numpy(), co, rtol=1e-7, atol=1e-9)
