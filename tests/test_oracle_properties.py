"""Property tests of the CPU oracle for the paths the reference's notebook does not pin (SURVEY.md §8c 'unpinned'):
the optimiser gradient against central finite differences of the objective, permutation invariance, and the
consistency of predict / log_predictive_density — for both kernels, with and without the mean function."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle.gpy_restated import GPRegressionOracle, logexp_f, logexp_finv


def make(kernel, mf, sfix, nfix, seed, m):
    rs = np.random.RandomState(seed)
    x = np.sort(rs.uniform(0, 8, m))
    y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
    return GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name=kernel, signal_variance=0.8,
                              signal_variance_fix=sfix, noise_variance=0.3, noise_variance_fix=nfix, mean_func=mf,
                              lengthscale=2.0, a_draw=0.4 if mf else None), x, y


@settings(max_examples=25, deadline=None)
@given(kernel=st.sampled_from(['rbf', 'linear']), mf=st.booleans(), sfix=st.booleans(), nfix=st.booleans(),
       seed=st.integers(0, 10 ** 6), m=st.integers(3, 40))
def test_objective_gradient_matches_finite_differences(kernel, mf, sfix, nfix, seed, m):
    o, _x, _y = make(kernel, mf, sfix, nfix, seed, m)
    t0 = o.optimizer_array().copy()
    if t0.size == 0:
        return
    f0, g0 = o.objective_grads(t0)
    for i in range(t0.size):
        h = 1e-5 * max(1.0, abs(t0[i]))
        tp, tm = t0.copy(), t0.copy()
        tp[i] += h
        tm[i] -= h
        fd = (o.objective_grads(tp)[0] - o.objective_grads(tm)[0]) / (2 * h)
        assert fd == pytest.approx(g0[i], rel=2e-5, abs=2e-6)


@settings(max_examples=20, deadline=None)
@given(kernel=st.sampled_from(['rbf', 'linear']), mf=st.booleans(), seed=st.integers(0, 10 ** 6), m=st.integers(2, 60))
def test_marginal_and_predictions_are_order_invariant(kernel, mf, seed, m):
    """The engine keeps cluster rows in arrival order; the reference re-extracts them in patient order. The GP
    does not care: LML, predictive mean and variance are invariant to the order of the data."""
    o, x, y = make(kernel, mf, True, False, seed, m)
    perm = np.random.RandomState(seed + 1).permutation(m)
    p = GPRegressionOracle(x[perm].reshape(-1, 1), y[perm].reshape(-1, 1), kernel_name=kernel, signal_variance=0.8,
                           noise_variance=0.3, mean_func=mf, lengthscale=2.0, a_draw=0.4 if mf else None)
    assert p.log_likelihood() == pytest.approx(o.log_likelihood(), rel=1e-9, abs=1e-9)
    xs = np.linspace(0, 9, 7).reshape(-1, 1)
    m0, v0 = o.predict(xs)
    m1, v1 = p.predict(xs)
    assert np.allclose(m0, m1, rtol=1e-8, atol=1e-10) and np.allclose(v0, v1, rtol=1e-8)


def test_logexp_roundtrip_and_limits():
    th = np.array([1e-8, 1e-3, 0.5, 4.0, 30.0, 40.0, 500.0])
    assert np.allclose(logexp_f(logexp_finv(th)), th, rtol=1e-12)
    assert logexp_f(np.array([-800.0]))[0] > 0.0
    assert logexp_f(np.array([50.0]))[0] == 50.0


def test_density_is_the_gaussian_of_predict():
    o, x, y = make('rbf', True, True, False, 5, 30)
    xs = np.array([[0.5], [3.0], [7.5]])
    ys = np.array([[0.2], [-0.4], [-1.1]])
    mu, var = o.predict(xs)
    lpd = o.log_predictive_density(xs, ys)
    ref = -0.5 * np.log(2 * np.pi * var) - 0.5 * (ys - mu) ** 2 / var
    assert np.allclose(lpd, ref, rtol=1e-12)


def test_rank_n_score_identities_against_refactorisation():
    """The identities behind the engine's rank-n score corrections (oracle/rank_updates.py):
    predictive moments of query visits after a block of the cluster's data leaves / joins, computed from the OLD factor
    and the cached T = M k, equal those of a refactorised GP on the changed data (GPRegressionOracle) to 1e-10."""
    from oracle import rank_updates
    from oracle.gpy_restated import GPRegressionOracle, GPY_JITTER
    rs = np.random.RandomState(3)
    for kernel in ('rbf', 'linear'):
        m, nq = 90, 7
        x = np.sort(rs.uniform(0, 8, m)); y = (40 - 5 * x + rs.randn(m) * 3 - 25) / 15
        xq = rs.uniform(0.2, 8, nq)
        gp = GPRegressionOracle(x, y, kernel_name=kernel, noise_variance=0.3, a_draw=0.4, lengthscale=2.5)
        L = gp._L
        M = np.linalg.inv(L)
        alpha = gp._alpha[:, 0]
        Kx, _ = gp._K(gp.X, xq.reshape(-1, 1))
        T_q = M @ Kx
        s_q = np.sum(T_q * T_q, axis=0)
        mu_q = Kx.T @ alpha                                       # without the mean function (as the engine caches it)
        kdiag = gp._Kdiag(xq.reshape(-1, 1))
        # --- a block leaves
        rows = np.arange(31, 40)
        keep = np.setdiff1d(np.arange(m), rows)
        s1, mu1 = rank_updates.leave_block(M, alpha, T_q, s_q, mu_q, rows)
        ref = GPRegressionOracle(x[keep], y[keep], kernel_name=kernel, noise_variance=0.3, a_draw=0.4, lengthscale=2.5)
        mu_ref, var_ref = ref._raw_predict(xq)
        assert np.allclose(kdiag - s1, var_ref[:, 0], rtol=1e-10, atol=1e-12)
        assert np.allclose(mu1 + ref._mean(xq.reshape(-1, 1))[:, 0], mu_ref[:, 0], rtol=1e-10, atol=1e-12)
        # --- a block joins
        xb = rs.uniform(0.3, 8, 6); yb = (40 - 5 * xb + rs.randn(6) * 3 - 25) / 15
        Kb, _ = gp._K(gp.X, xb.reshape(-1, 1))
        T_B = M @ Kb
        Kbb, _ = gp._K(xb.reshape(-1, 1))
        S_BB = Kbb + (gp.noise + GPY_JITTER) * np.eye(6) - T_B.T @ T_B
        mu_b = Kb.T @ alpha + gp._mean(xb.reshape(-1, 1))[:, 0]
        K_Bq, _ = gp._K(xb.reshape(-1, 1), xq.reshape(-1, 1))
        s2, mu2 = rank_updates.join_block(T_q, s_q, mu_q, T_B, K_Bq, S_BB, yb - mu_b)
        ref = GPRegressionOracle(np.concatenate([x, xb]), np.concatenate([y, yb]), kernel_name=kernel, noise_variance=0.3,
                                 a_draw=0.4, lengthscale=2.5)
        mu_ref, var_ref = ref._raw_predict(xq)
        assert np.allclose(kdiag - s2, var_ref[:, 0], rtol=1e-10, atol=1e-12)
        assert np.allclose(mu2 + ref._mean(xq.reshape(-1, 1))[:, 0], mu_ref[:, 0], rtol=1e-10, atol=1e-12)
        # --- alpha of the changed clusters in closed form (k_append_finalize / k_rm_alpha): no pass over the factor
        assert np.allclose(rank_updates.alpha_after_join(M, alpha, T_B, S_BB, yb - mu_b), ref._alpha[:, 0], rtol=1e-9, atol=1e-11)
        ref = GPRegressionOracle(x[keep], y[keep], kernel_name=kernel, noise_variance=0.3, a_draw=0.4, lengthscale=2.5)
        assert np.allclose(rank_updates.alpha_after_leave(M, alpha, rows), ref._alpha[:, 0], rtol=1e-9, atol=1e-11)


def test_repeated_inputs_collapse_identity():
    """The identity the native refit's shadow clusters rest on (mogp_b200/csrc/refit.inl): n observations at one input are one
    observation of their mean with noise s / n, times (2 pi s)^-(n-1)/2 n^-1/2 exp(-SS / 2s), s = noise + 1e-8 (GPy's diagonal
    jitter rides with the noise).  LML and d LML / d noise of the oracle on the full data = weighted GP on the merged data +
    the closed-form term."""
    rs = np.random.RandomState(5)
    m = 90
    x = np.sort(rs.uniform(0, 8, m))
    x[:20] = 0.0
    x[40:43] = x[40]
    y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
    o = GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name='rbf', signal_variance=0.8,
                           signal_variance_fix=False, noise_variance=0.3, noise_variance_fix=False, mean_func=False,
                           lengthscale=2.0, a_draw=None)
    lml_full = o.log_likelihood()
    d_noise_full = o._grad_lml[-1]
    xu, inv, cnt = np.unique(x, return_inverse=True, return_counts=True)
    ybar = np.bincount(inv, weights=y) / cnt
    ss = np.sum((y - ybar[inv]) ** 2)
    s = 0.3 + 1e-8
    K = 0.8 * np.exp(-0.5 * (xu[:, None] - xu[None, :]) ** 2 / 2.0 ** 2)

    def collapsed(s):
        Ky = K + np.diag(s / cnt)
        L = np.linalg.cholesky(Ky)
        a = np.linalg.solve(Ky, ybar)
        lml = -0.5 * len(xu) * np.log(2 * np.pi) - np.sum(np.log(np.diag(L))) - 0.5 * ybar @ a
        n1 = np.sum(cnt - 1.0)
        corr = -(0.5 * n1 * np.log(2 * np.pi * s) + 0.5 * np.sum(np.log(cnt)) + 0.5 * ss / s)
        Ki = np.linalg.inv(Ky)
        dl = 0.5 * (np.outer(a, a) - Ki)
        d = np.sum(np.diag(dl) / cnt) - (0.5 * n1 / s - 0.5 * ss / s ** 2)
        return lml + corr, d

    lml_c, d_c = collapsed(s)
    assert lml_c == pytest.approx(lml_full, rel=1e-11)
    assert d_c == pytest.approx(d_noise_full, rel=1e-9)
