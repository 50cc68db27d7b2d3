"""Pins the CPU oracle against the known-answer tests the reference's own test-suite holds for
this path (the reference itself cannot be imported here: TensorFlow is absent).  Each test
names the reference test it restates."""
import numpy as np
import pytest
import scipy.stats
from numpy.testing import assert_allclose

from oracle import gp_oracle as O


def test_rbf_vs_loop_reference():
    """tests/gpflow/kernels/test_kernels.py:94-101 with tests/gpflow/kernels/reference.py:13-27."""
    rng = np.random.RandomState(1)
    X = rng.randn(3, 1)
    var, ell = 2.3, 1.4
    K = O.SquaredExponential(variance=var, lengthscales=ell)(X)
    ref = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            d = X[i] - X[j]
            ref[i, j] = var * np.exp(-0.5 * d.dot(d) / ell ** 2)
    assert_allclose(K, ref)


@pytest.mark.parametrize("cls", [O.SquaredExponential, O.Matern12, O.Matern32, O.Matern52,
                                 O.RationalQuadratic, O.Exponential, O.Linear, O.Constant, O.White])
def test_kernel_symmetry_and_diag(cls):
    """test_kernels.py:262-265 (K(X)==K(X,X) except White) and :322-326 (diag(K)==K_diag)."""
    rng = np.random.RandomState(0)
    X = rng.randn(7, 3)
    k = cls()
    if cls is not O.White:
        assert_allclose(k(X), k(X, X), atol=1e-14)
    else:  # test_kernels.py:365-375
        assert not np.allclose(k(X), k(X, X))
    assert_allclose(np.diag(k(X)), k(X, full_cov=False), atol=1e-14)


def test_rq_limit_is_rbf():
    """test_kernels.py:105-113."""
    rng = np.random.RandomState(1)
    X = rng.randn(6, 2)
    assert_allclose(O.RationalQuadratic(alpha=1e8)(X), O.SquaredExponential()(X), atol=1e-7)


def test_sum_product_active_dims():
    """test_kernels.py:349-361, 425-429, 433-455."""
    rng = np.random.RandomState(3)
    X = rng.randn(8, 3)
    k1, k2 = O.Matern32(active_dims=[0, 1]), O.SquaredExponential(active_dims=[2])
    assert_allclose((k1 + k2)(X), k1(X) + k2(X))
    assert_allclose((k1 * k2)(X), k1(X) * k2(X))
    # product of 1-D RBFs over separate dims == ARD RBF
    ells = np.array([0.7, 1.3, 2.1])
    prod = O.Product([O.SquaredExponential(lengthscales=ells[i], active_dims=[i]) for i in range(3)])
    assert_allclose(prod(X), O.SquaredExponential(lengthscales=ells)(X), atol=1e-13)
    # nested same-class combos are flattened (base.py:246-254)
    assert len(((k1 + k2) + k1).kernels) == 3 and len(((k1 + k2) * k1).kernels) == 2


def test_matern_finite_near_zero_distance():
    """kernels/test_scaled_euclid_dist.py:40-57 — the 1e-36 clip keeps sqrt finite."""
    X = np.random.RandomState(0).randn(100, 100)
    for cls in (O.Matern12, O.Matern32, O.Matern52, O.Exponential):
        assert np.all(np.isfinite(cls()(X)))


def test_psd():
    """kernels/test_positive_semidefinite.py:35-51."""
    X = np.random.RandomState(0).randn(50, 3)
    for cls in (O.SquaredExponential, O.Matern12, O.Matern32, O.Matern52, O.Linear):
        assert np.linalg.eigvalsh(cls()(X)).min() > -1e-12


@pytest.mark.parametrize("ncol_x,ncol_mu", [(10, 10), (1, 10), (1, 1)])
@pytest.mark.parametrize("eye", [False, True])
def test_multivariate_normal_vs_scipy(ncol_x, ncol_mu, eye):
    """tests/gpflow/test_logdensities.py:113-128."""
    rng = np.random.RandomState(0)
    x, mu = rng.randn(4, ncol_x), rng.randn(4, ncol_mu)
    cs = np.eye(4) if eye else rng.randn(4, 4)
    cov = cs @ cs.T
    L = np.linalg.cholesky(cov)
    got = O.multivariate_normal(x, mu, L)
    want = [scipy.stats.multivariate_normal.logpdf(x[:, i if ncol_x > 1 else 0], mu[:, i], cov)
            for i in range(ncol_mu)]
    assert_allclose(got, want)


def _make_kl_datum():  # tests/gpflow/test_kullback_leiblers.py:106-119
    from types import SimpleNamespace
    rng = np.random.RandomState(0)
    M, N = 5, 4
    mu = rng.randn(M, N)
    A = rng.randn(M, M)
    K = A @ A.T + 1e-6 * np.eye(M)
    sqrt = np.array([np.tril(rng.randn(M, M)) for _ in range(N)])
    sqrt_diag = rng.randn(M, N)
    Kb = rng.randn(N, M, M)
    K_batch = 0.1 * (Kb + Kb.transpose(0, 2, 1)) + np.eye(M)[None]
    return SimpleNamespace(M=M, N=N, mu=mu, K=K, sqrt=sqrt, sqrt_diag=sqrt_diag, K_batch=K_batch,
                           K_cholesky=np.linalg.cholesky(K))


KLDatum = _make_kl_datum()


def _kl_1d(q_mu, q_sigma, p_var=1.0):  # test_kullback_leiblers.py:94-98
    q_var = np.square(q_sigma)
    return np.sum(0.5 * (q_var / p_var + np.square(q_mu) / p_var - 1 + np.log(p_var / q_var)))


@pytest.mark.parametrize("white", [True, False])
def test_kl_oned(white):
    """test_kullback_leiblers.py:215-229."""
    rng = np.random.RandomState(0)
    mu1d, s1d = rng.randn(1, 1), rng.rand(1, 1) + 0.1
    K1d = rng.rand(1, 1) + 0.1
    kl = O.gauss_kl(mu1d, s1d, None if white else K1d)
    assert_allclose(kl, _kl_1d(mu1d, s1d, 1.0 if white else K1d))
    kl = O.gauss_kl(mu1d, s1d[None], None if white else K1d)
    assert_allclose(kl, _kl_1d(mu1d, s1d, 1.0 if white else K1d))


def test_kl_invariants():
    """test_kullback_leiblers.py:122-131 (K vs K_cholesky), :135-147 (diag vs dense),
    :151-164 (K=I vs white), :168-191 (batch sums)."""
    D = KLDatum
    for qs in (D.sqrt, D.sqrt_diag):
        assert_allclose(O.gauss_kl(D.mu, qs, D.K), O.gauss_kl(D.mu, qs, K_cholesky=D.K_cholesky))
        assert_allclose(O.gauss_kl(D.mu, qs, np.eye(D.M)), O.gauss_kl(D.mu, qs, None), atol=1e-10)
    dense_from_diag = np.array([np.diag(D.sqrt_diag[:, i]) for i in range(D.N)])
    for K in (None, D.K, D.K_batch):
        assert_allclose(O.gauss_kl(D.mu, D.sqrt_diag, K), O.gauss_kl(D.mu, dense_from_diag, K))
    total = O.gauss_kl(D.mu, D.sqrt, D.K_batch)
    parts = sum(O.gauss_kl(D.mu[:, i:i + 1], D.sqrt[i:i + 1], D.K_batch[i]) for i in range(D.N))
    assert_allclose(total, parts)
    total = O.gauss_kl(D.mu, D.sqrt, D.K)
    parts = sum(O.gauss_kl(D.mu[:, i:i + 1], D.sqrt[i:i + 1], D.K) for i in range(D.N))
    assert_allclose(total, parts)


@pytest.mark.parametrize("full_cov", [True, False])
def test_base_conditional_vs_explicit_inverse(full_cov):
    """tests/gpflow/conditionals/test_conditionals.py:168-214."""
    rng = np.random.RandomState(123)
    Dy, N, M, Dx = 5, 4, 3, 2
    X, Z = rng.randn(N, Dx), rng.randn(M, Dx)
    kern = O.Matern52(lengthscales=0.5)
    q_mu = rng.randn(M, Dy)
    q_sqrt = np.tril(rng.randn(Dy, M, M), -1)
    Kmm = kern(Z, Z) + np.eye(M) * 1e-6
    Kmn, Knn = kern(Z, X), kern(X, X)
    S = q_sqrt @ q_sqrt.transpose(0, 2, 1)
    Ki = np.linalg.inv(Kmm)
    mean_np = Kmn.T @ Ki @ q_mu
    cov_np = Knn[None] + Kmn.T[None] @ Ki[None] @ (S - Kmm[None]) @ Ki[None] @ Kmn[None]
    mean, cov = O.svgp_predict_f(X, Z, kern, q_mu, q_sqrt, whiten=False, full_cov=full_cov)
    if not full_cov:
        cov_np = np.diagonal(cov_np, axis1=-1, axis2=-2).T
    assert_allclose(mean, mean_np, rtol=1e-6, atol=1e-9)
    assert_allclose(cov, cov_np, rtol=1e-6, atol=1e-9)


def test_whiten_vs_unwhitened_conditional():
    """test_conditionals.py:87-102: conditional(V=L^-1 F, white) == conditional(F)."""
    rng = np.random.RandomState(123)
    Nn, Mn, Ln = 10, 20, 2
    k = O.Matern32() + O.White(variance=0.01)
    Xs, X = rng.randn(Nn, 1), rng.randn(Mn, 1)
    F = rng.randn(Mn, Ln)
    Kmm = k(X) + 1e-6 * np.eye(Mn)
    Lm = np.linalg.cholesky(Kmm)
    V = np.linalg.solve(Lm, F)
    m1, v1 = O.base_conditional(k(X, Xs), Kmm, k(Xs, full_cov=False), F, white=False)
    m2, v2 = O.base_conditional(k(X, Xs), Kmm, k(Xs, full_cov=False), V, white=True)
    assert_allclose(m1, m2, atol=1e-8)
    assert_allclose(v1, v2, atol=1e-8)


def test_diag_vs_chol_q_sqrt():
    """test_conditionals.py:68-84."""
    rng = np.random.RandomState(123)
    Nn, Mn, Ln = 10, 20, 2
    k = O.Matern32() + O.White(variance=0.01)
    Xs, Z = rng.randn(Nn, 1), rng.randn(Mn, 1)
    mu = rng.randn(Mn, Ln)
    sd = rng.rand(Mn, Ln)
    chol = np.array([np.diag(sd[:, i]) for i in range(Ln)])
    m1, v1 = O.svgp_predict_f(Xs, Z, k, mu, sd, whiten=False)
    m2, v2 = O.svgp_predict_f(Xs, Z, k, mu, chol, whiten=False)
    assert_allclose(m1, m2)
    assert_allclose(v1, v2)


class EqDatum:  # tests/integration/test_method_equivalence.py:30-40
    rng = np.random.RandomState(0)
    X = rng.rand(20, 1) * 10
    Y = np.sin(X) + 0.9 * np.cos(X * 1.6) + rng.randn(*X.shape) * 0.8
    Y = np.tile(Y, 2)
    Xtest = rng.rand(10, 1) * 10


def test_method_equivalence_fixed_hyperparameters():
    """tests/integration/test_method_equivalence.py:181-241 at fixed hyper-parameters: with Z=X
    and a Gaussian likelihood SGPR's bound equals the GPR LML (up to the Kuu jitter), and SVGP
    with q(u) from `compute_qu` attains the SGPR bound and the same predictions."""
    D = EqDatum
    k = O.SquaredExponential(variance=1.3, lengthscales=1.7)
    s2 = 0.4
    lml = O.gpr_log_marginal_likelihood(D.X, D.Y, k, s2)
    elbo = O.sgpr_elbo(D.X, D.Y, k, D.X.copy(), s2)
    assert_allclose(elbo, lml, rtol=1e-5)
    assert elbo <= lml + 1e-9
    mu, cov = O.sgpr_compute_qu(D.X, D.Y, k, D.X.copy(), s2)
    q_sqrt = np.tile(np.linalg.cholesky(cov + 1e-12 * np.eye(20))[None], (2, 1, 1))
    e2 = O.svgp_elbo(D.X, D.Y, D.X.copy(), k, mu, q_sqrt, s2, whiten=False)
    assert_allclose(e2, elbo, rtol=1e-4)
    m_g, v_g = O.gpr_predict_f(D.X, D.Y, k, s2, D.Xtest)
    m_s, v_s = O.sgpr_predict_f(D.X, D.Y, k, D.X.copy(), s2, D.Xtest)
    m_v, v_v = O.svgp_predict_f(D.Xtest, D.X.copy(), k, mu, q_sqrt, whiten=False)
    assert_allclose(m_s, m_g, rtol=1e-3, atol=1e-4)
    assert_allclose(v_s, v_g, rtol=1e-3, atol=1e-4)
    assert_allclose(m_v, m_g, rtol=1e-3, atol=1e-4)
    assert_allclose(v_v, v_g, rtol=1e-3, atol=1e-4)


def test_sgpr_qu_equals_predict_at_Z():
    """tests/gpflow/models/test_sgpr.py:29-44 (at fixed hyper-parameters)."""
    rng = np.random.RandomState(0)
    X, _, Z = rng.randn(100, 2), rng.randn(100, 1), rng.randn(20, 2)
    rng1 = np.random.RandomState(1)
    Y = np.sin(X @ np.array([[-1.4], [0.5]])) + 0.5 * rng1.randn(100, 1)
    k = O.SquaredExponential()
    # The identity is exact only as jitter -> 0 (predict_f(Z) uses the un-jittered Kus); the
    # reference test reaches 1e-5 at optimised hyper-parameters, here we shrink the jitter.
    mu, cov = O.sgpr_compute_qu(X, Y, k, Z, 1.0, jitter=1e-10)
    m, v = O.sgpr_predict_f(X, Y, k, Z, 1.0, Z, full_cov=True, jitter=1e-10)
    assert_allclose(mu, m, rtol=1e-5, atol=1e-5)
    assert_allclose(cov[None], v, rtol=1e-5, atol=1e-5)


def test_svgp_qdiag_equals_full():
    """tests/gpflow/models/test_svgp.py:60-129: diagonal q_sqrt ELBO == dense diag q_sqrt ELBO."""
    rng = np.random.RandomState(0)
    X, Y, Z = rng.randn(30, 2), rng.randn(30, 2), rng.randn(7, 2)
    k = O.SquaredExponential() + O.White(variance=0.1)
    q_mu = rng.randn(7, 2)
    qd = rng.rand(7, 2) + 0.2
    qf = np.array([np.diag(qd[:, i]) for i in range(2)])
    for white in (True, False):
        a = O.svgp_elbo(X, Y, Z, k, q_mu, qd, 0.3, whiten=white, num_data=100)
        b = O.svgp_elbo(X, Y, Z, k, q_mu, qf, 0.3, whiten=white, num_data=100)
        assert_allclose(a, b)


def test_cached_posterior_equals_fused():
    """tests/gpflow/posteriors/test_posteriors.py (fused vs precomputed), posteriors.py:694-822."""
    rng = np.random.RandomState(5)
    Z, Xn = rng.randn(9, 2), rng.randn(11, 2)
    k = O.Matern52(lengthscales=0.8)
    q_mu = rng.randn(9, 3)
    q_sqrt = np.array([np.tril(rng.randn(9, 9)) * 0.3 + np.eye(9) for _ in range(3)])
    for white in (True, False):
        m1, v1 = O.svgp_predict_f(Xn, Z, k, q_mu, q_sqrt, whiten=white)
        alpha, Qinv = O.svgp_cached_alpha_qinv(Z, k, q_mu, q_sqrt, whiten=white)
        m2, v2 = O.svgp_predict_f_cached(Xn, Z, k, alpha, Qinv)
        assert_allclose(m1, m2, rtol=1e-8, atol=1e-9)
        assert_allclose(v1, v2, rtol=1e-6, atol=1e-8)


def test_log_density_by_hand():
    """tests/gpflow/models/test_model_predict.py:105-135."""
    rng = np.random.RandomState(2)
    mu, var, Y = rng.randn(6, 2), rng.rand(6, 2) + 0.1, rng.randn(6, 2)
    s2 = 0.3
    got = O.gaussian_predict_log_density(mu, var, Y, s2)
    want = scipy.stats.norm.logpdf(Y, loc=mu, scale=np.sqrt(var + s2)).sum(-1)
    assert_allclose(got, want)


# ---- sibling models on the same operators ------------------------------------------------------
def test_sgpr_upper_bound_brackets_the_marginal_likelihood():
    """tests/integration/test_method_equivalence.py:297-327 (DatumUpper: rng 123, X rand(100,1), offset 5.3) at fixed
    hyper-parameters: elbo < GPR lml < upper_bound."""
    rng = np.random.default_rng(123)
    X = rng.random((100, 1))
    Y = np.sin(1.5 * 2 * np.pi * X) + rng.standard_normal(X.shape) * 0.1 + 5.3
    assert Y.mean() > 5.0
    k = O.SquaredExponential(variance=1.3, lengthscales=0.3)
    mf = O.ConstantMean([5.0])
    Z = X[:10].copy()
    elbo = O.sgpr_elbo(X, Y, k, Z, 0.05, mf)
    lml = O.gpr_log_marginal_likelihood(X, Y, k, 0.05, mf)
    ub = O.sgpr_upper_bound(X, Y, k, Z, 0.05, mf)
    assert elbo < lml < ub
    # with Z = X both bounds are tight (the trace term vanishes)
    assert abs(O.sgpr_upper_bound(X, Y, k, X.copy(), 0.05, mf, jitter=1e-10) - lml) < 1e-4 * abs(lml)


def test_vgp_with_exact_posterior_equals_gpr():
    """VGP (vgp.py:111-161) with q(v) set to the exact whitened posterior: ELBO == GPR log marginal likelihood and
    the predictions coincide -- the fixed point the reference's method-equivalence test reaches by optimisation
    (tests/integration/test_method_equivalence.py:181-241)."""
    rng = np.random.default_rng(7)
    N = 30
    X = rng.standard_normal((N, 2))
    Y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    k = O.Matern32(variance=0.9, lengthscales=1.1)
    s2, jit = 0.2, 1e-10
    K = k(X) + jit * np.eye(N)
    L = np.linalg.cholesky(K)
    Ky = K + s2 * np.eye(N)
    mu = K @ np.linalg.solve(Ky, Y)
    S = K - K @ np.linalg.solve(Ky, K)
    q_mu = np.linalg.solve(L, mu)
    q_sqrt = np.linalg.solve(L, np.linalg.cholesky(S + 1e-14 * np.eye(N)))[None]
    assert abs(O.vgp_elbo(X, Y, k, q_mu, q_sqrt, s2, jitter=jit) - O.gpr_log_marginal_likelihood(X, Y, k, s2)) < 1e-6
    Xn = rng.standard_normal((5, 2))
    m1, v1 = O.vgp_predict_f(X, k, q_mu, q_sqrt, Xn, jitter=jit)
    m2, v2 = O.gpr_predict_f(X, Y, k, s2, Xn)
    np.testing.assert_allclose(m1, m2, atol=1e-8)
    np.testing.assert_allclose(v1, v2, atol=1e-8)


# ---- further reference tests restated on the oracle (tests/gpflow/kernels/test_kernels.py) ----------------------
def test_white_sym_differs_from_explicit_x2():
    """test_kernels.py:365-376: White()(X) != White()(X, X) (statics.py:57-63)."""
    X = np.random.RandomState(1).randn(10, 3)
    k = O.White()
    assert not np.allclose(k(X), k(X, X))
    assert np.all(k(X, X) == 0.0) and np.allclose(k(X), np.eye(10))


_SLICE_CLASSES = [O.SquaredExponential, O.RationalQuadratic, O.Exponential, O.Matern12, O.Matern32, O.Matern52,
                  O.Constant, O.Linear, O.Polynomial]


@pytest.mark.parametrize("cls", _SLICE_CLASSES)
def test_slice_symmetric_and_asymmetric(cls):
    """test_kernels.py:396-424: active_dims=[i] on X equals the same kernel on the sliced column, list and slice forms."""
    rng = np.random.RandomState(1)
    X, Z = rng.randn(20, 2), rng.randn(12, 2)
    k1, k2, k3 = cls(active_dims=[0]), cls(active_dims=[1]), cls(active_dims=slice(0, 1))
    assert np.allclose(k1(X), k3(X[:, :1])) and np.allclose(k2(X), k3(X[:, 1:]))
    assert np.allclose(k1(X, Z), k3(X[:, :1], Z[:, :1])) and np.allclose(k2(X, Z), k3(X[:, 1:], Z[:, 1:]))


def test_product_and_active_product():
    """test_kernels.py:425-458."""
    rng = np.random.RandomState(1)
    X = rng.randn(30, 2)
    a, b = O.Matern32(), O.Matern52(lengthscales=0.3)
    assert np.allclose(a(X) * b(X), (O.Matern32() * O.Matern52(lengthscales=0.3))(X))
    for N, D in ((30, 4), (10, 7)):
        X = rng.randn(N, D)
        dims, idx, ls = list(range(D)), int(rng.randint(0, D)), rng.uniform(1.0, 7.0, D)
        k_rest = O.SquaredExponential(lengthscales=np.hstack([ls[:idx], ls[idx + 1:]]), active_dims=dims[:idx] + dims[idx + 1:])
        k_one = O.SquaredExponential(lengthscales=ls[idx], active_dims=[idx])
        k_all = O.SquaredExponential(lengthscales=ls, active_dims=dims)
        assert np.allclose(k_all(X), (k_rest * k_one)(X))


def test_kernel_call_diag_and_x2_errors():
    """test_kernels.py:621-627: full_cov=False with an explicit X2 is a ValueError (base.py:203-204)."""
    rng = np.random.RandomState(1)
    X, X2 = rng.randn(4, 1), rng.randn(5, 1)
    for k in (O.SquaredExponential(), O.Matern32() + O.White(), O.Linear() * O.Constant()):
        with pytest.raises(ValueError):
            k(X, X2, full_cov=False)


def test_diag_equals_diagonal_of_full():
    """test_kernels.py:322-327 (test_diags)."""
    X = np.random.RandomState(1).randn(15, 3)
    for k in (O.SquaredExponential(0.7, 1.3), O.Matern12(), O.Linear(0.4), O.Polynomial(2.0, 0.5, 0.9), O.Constant(0.3), O.White(0.2),
              O.Matern52() * O.Linear() + O.White(0.1)):
        np.testing.assert_allclose(k(X, full_cov=False), np.diag(k(X)), rtol=1e-6)  # Matern12: sqrt(eps) noise on the diagonal (ops.py:109-111)


# ---- tests/gpflow/test_kullback_leiblers.py, conditionals/test_conditionals.py -------------------------------------
@pytest.mark.parametrize("shared_k", [True, False])
@pytest.mark.parametrize("diag", [True, False])
def test_sumkl_equals_batchkl(shared_k, diag):
    """test_kullback_leiblers.py:166-192: gauss_kl sums the per-column KLs (kullback_leiblers.py:72-74)."""
    rng = np.random.RandomState(0)
    M, N = 5, 4
    mu = rng.randn(M, N)
    sqrt = np.stack([np.tril(rng.randn(M, M)) for _ in range(N)])
    sqrt_diag = rng.rand(M, N) + 0.1
    A = rng.randn(M, M)
    K = A @ A.T + 1e-3 * np.eye(M)
    K_batch = np.stack([K * (1.0 + 0.1 * n) for n in range(N)])
    s = sqrt_diag if diag else sqrt
    kl_batch = O.gauss_kl(mu, s, K if shared_k else K_batch)
    kl_sum = 0.0
    for n in range(N):
        s_n = sqrt_diag[:, n][:, None] if diag else sqrt[n][None]
        K_n = K if shared_k else K_batch[n][None]
        kl_sum += O.gauss_kl(mu[:, n][:, None], s_n, K=K_n)
    assert abs(kl_sum - kl_batch) < 1e-9 * max(1.0, abs(kl_batch))


@pytest.mark.parametrize("white", [True, False])
def test_conditional_diag_q_sqrt_equals_cholesky_q_sqrt(white):
    """conditionals/test_conditionals.py:68-84 (test_diag) on the fixture of :28-66 (RandomState(123), Nn=10, Mn=20,
    Ln=2, Matern32 + White(0.01))."""
    rng = np.random.RandomState(123)
    Nn, Mn, Ln = 10, 20, 2
    Xs, X = rng.rand(Nn, 1), rng.rand(Mn, 1)
    k = O.Matern32() + O.White(0.01)
    mu = rng.rand(Mn, Ln)
    sqrt = rng.rand(Mn, Ln)
    chol = np.stack([np.diag(sqrt[:, i]) for i in range(Ln)])
    Kmm = k(X) + 1e-6 * np.eye(Mn)
    Kmn, Knn = k(X, Xs), k(Xs, full_cov=False)
    m1, v1 = O.base_conditional(Kmn, Kmm, Knn, mu, q_sqrt=sqrt, white=white)
    m2, v2 = O.base_conditional(Kmn, Kmm, Knn, mu, q_sqrt=chol, white=white)
    np.testing.assert_allclose(m1, m2, atol=1e-12)
    np.testing.assert_allclose(v1, v2, atol=1e-12)


# ---- kernels widened in round 2 (restating the reference's own checks) -----------------------------------------------
def test_periodic_reduces_to_base_on_mapped_inputs_and_cosine_identity():
    """gpflow/kernels/periodic.py docstring: the periodic kernel is the base kernel on u = (cos x, sin x); for the
    SquaredExponential base sum_d sin^2(pi (x - x') / p) / l^2 = |u - u'|^2 / (4 l^2) with u at angle 2 pi x / p.
    Cosine: cos(2 pi (a - b)) = cos 2 pi a cos 2 pi b + sin 2 pi a sin 2 pi b (a rank-2 kernel, so PSD)."""
    rng = np.random.default_rng(4)
    X = rng.standard_normal((9, 2))
    p, ell = np.array([1.5, 0.7]), 0.9
    kper = O.Periodic(O.SquaredExponential(1.3, ell), p)
    ang = 2 * np.pi * X / p
    U = np.concatenate([np.cos(ang), np.sin(ang)], axis=1)
    d2 = ((U[:, None, :] - U[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(kper(X), 1.3 * np.exp(-0.5 * d2 / (4 * ell ** 2)), rtol=1e-12)
    kc = O.Cosine(0.8, [0.5, 2.0])
    a = (X / np.array([0.5, 2.0])).sum(1)
    ref = 0.8 * (np.outer(np.cos(2 * np.pi * a), np.cos(2 * np.pi * a)) + np.outer(np.sin(2 * np.pi * a), np.sin(2 * np.pi * a)))
    np.testing.assert_allclose(kc(X), ref, rtol=1e-10, atol=1e-12)


def test_arccosine_changepoints_coregion_symmetric_psd_and_diag():
    """tests/gpflow/kernels/test_kernels.py: symmetric, PSD, K_diag == diag(K) (ArcCosine orders 0-2;
    ChangePoints with steep sigmoids switches between its regimes; Coregion indexes B = W W^T + diag(kappa))."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((20, 3))
    for order in (0, 1, 2):
        k = O.ArcCosine(order, 1.1, [0.5, 1.5, 1.0], 0.7)
        Km = k(X)
        np.testing.assert_allclose(Km, Km.T, rtol=1e-12)
        assert np.linalg.eigvalsh(Km).min() > -1e-8
        np.testing.assert_allclose(np.diag(Km), k(X, full_cov=False), rtol=1e-6)  # acos near 1 (jitter 1e-15)
    x = np.linspace(0, 1, 30)[:, None]
    cp = O.ChangePoints([O.SquaredExponential(1.0, 0.3), O.Matern32(2.0, 0.2)], [0.5], 1e4)
    Km = cp(x)
    lo, hi = x[:, 0] < 0.49, x[:, 0] > 0.51
    np.testing.assert_allclose(Km[np.ix_(lo, lo)], O.SquaredExponential(1.0, 0.3)(x[lo]), atol=1e-9)
    np.testing.assert_allclose(Km[np.ix_(hi, hi)], O.Matern32(2.0, 0.2)(x[hi]), atol=1e-9)
    np.testing.assert_allclose(Km[np.ix_(lo, hi)], 0.0, atol=1e-9)
    W = rng.standard_normal((4, 2))
    cg = O.Coregion(4, 2, W=W, kappa=np.array([1.0, 2.0, 0.5, 0.1]))
    idx = np.array([[3.0], [0.0], [3.0], [1.0]])
    B = W @ W.T + np.diag([1.0, 2.0, 0.5, 0.1])
    np.testing.assert_allclose(cg(idx), B[np.ix_([3, 0, 3, 1], [3, 0, 3, 1])], rtol=1e-14)


def test_fitc_equals_gpr_when_inducing_points_are_the_data_and_mixing_identities():
    """tests/gpflow/models/test_sgpr.py / test_method_equivalence.py idea: with Z = X the FITC approximation is exact
    (nu = sigma^2 up to the jitter), so its LML and predictions equal GPR's.  mix_latent_gp with W = I is the identity,
    and its full covariance contracts to the marginal forms."""
    rng = np.random.default_rng(6)
    X = rng.standard_normal((30, 2))
    Y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((30, 2))
    k = O.Matern32(1.2, 0.9)
    np.testing.assert_allclose(O.gprfitc_lml(X, Y, k, X, 0.2), O.gpr_log_marginal_likelihood(X, Y, k, 0.2), rtol=1e-5)
    Xs = rng.standard_normal((7, 2))
    mf, vf = O.gprfitc_predict_f(X, Y, k, X, 0.2, Xs)
    mg, vg = O.gpr_predict_f(X, Y, k, 0.2, Xs)
    np.testing.assert_allclose(mf, mg, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(vf, vg, rtol=1e-4, atol=1e-6)
    W = rng.standard_normal((3, 2))
    gm, gv = rng.standard_normal((5, 2)), rng.uniform(0.1, 1, (5, 2))
    m1, v1 = O.mix_latent_gp(W, gm, gv, False, False)
    m2, v2 = O.mix_latent_gp(W, gm, gv, False, True)
    np.testing.assert_allclose(np.einsum("npp->np", v2), v1, rtol=1e-12)
    A = rng.standard_normal((2, 5, 5))
    gcov = A @ np.transpose(A, (0, 2, 1))
    _, c1 = O.mix_latent_gp(W, gm, gcov, True, False)
    _, c2 = O.mix_latent_gp(W, gm, gcov, True, True)
    np.testing.assert_allclose(np.einsum("npmp->pnm", c2), c1, rtol=1e-12)
    m3, v3 = O.mix_latent_gp(np.eye(2), gm, gv, False, False)
    np.testing.assert_allclose(m3, gm), np.testing.assert_allclose(v3, gv)


def test_sample_mvn_moments_and_shapes():
    """tests/gpflow/conditionals/test_conditionals.py::test_sample_mvn: sample mean / covariance converge to the inputs."""
    rng = np.random.default_rng(8)
    mean = np.array([[1.0, -2.0]])
    A = np.array([[1.0, 0.0], [0.8, 0.6]])
    cov = (A @ A.T)[None]
    S = 200000
    s = O.sample_mvn(mean, cov, True, rng.standard_normal((1, 2, S)), jitter=0.0)       # [S, 1, 2]
    np.testing.assert_allclose(s.mean(0), mean, atol=1e-2)
    np.testing.assert_allclose(np.cov(s[:, 0, :].T), cov[0], atol=2e-2)
    d = O.sample_mvn(mean, np.array([[0.25, 4.0]]), False, rng.standard_normal((S, 1, 2)))
    np.testing.assert_allclose(d.std(0), [[0.5, 2.0]], rtol=2e-2)
