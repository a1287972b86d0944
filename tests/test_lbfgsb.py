"""The engine's L-BFGS-B state machine (mogp_b200/csrc/lbfgsb.hpp) against scipy.optimize.fmin_l_bfgs_b — the optimiser
paramz runs behind `model.optimize()` (mogp/obsmodel.py:94): same evaluation points in the same order (the count is
what a refit costs), same optimum.  CPU only: the state machine is host code behind the C ABI."""
import numpy as np
import pytest
from scipy.optimize import fmin_l_bfgs_b

from mogp_b200._capi import lbfgsb_minimize
from oracle.gpy_restated import GPRegressionOracle


def trace_scipy(fun, x0):
    pts = []

    def wrapped(x):
        pts.append(np.array(x, dtype=float))
        return fun(x)
    x, f, d = fmin_l_bfgs_b(wrapped, np.array(x0, dtype=float), maxfun=1000, maxiter=1000)
    return x, f, d, pts


def rosen(x):
    f = np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)
    g = np.zeros_like(x)
    g[:-1] += -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
    g[1:] += 200 * (x[1:] - x[:-1] ** 2)
    return f, g


def quad(scales):
    A = np.diag(scales)

    def fun(x):
        return 0.5 * x @ A @ x + np.sum(np.cos(x)), A @ x - np.sin(x)
    return fun


def compare(fun, x0, min_evals=3):
    xs, fs, ds, ps = trace_scipy(fun, x0)
    xe, fe, info = lbfgsb_minimize(fun, x0)
    pe = info['points']
    n = min(len(ps), len(pe))
    first_bad = next((i for i in range(n) if not np.allclose(ps[i], pe[i], rtol=1e-7, atol=1e-9)), None)
    assert first_bad is None, "evaluation {} differs: scipy {} engine {}".format(first_bad, ps[first_bad], pe[first_bad])
    assert len(pe) == len(ps) == ds['funcalls'], (len(pe), len(ps), ds['funcalls'])
    assert info['nit'] == ds['nit']
    assert np.allclose(xe, xs, rtol=1e-7, atol=1e-9) and abs(fe - fs) <= 1e-9 * max(1.0, abs(fs))
    assert len(ps) >= min_evals
    # same stopping reason: 0 = converged in SciPy's warnflag; 3 / 4 = the two convergence tests here
    assert (ds['warnflag'] == 0) == (info['task'] in (3, 4))
    return len(ps)


@pytest.mark.parametrize("n,seed", [(2, 0), (3, 1), (3, 2), (4, 3), (6, 4)])
def test_rosenbrock_same_evaluations_as_scipy(n, seed):
    rs = np.random.RandomState(seed)
    compare(rosen, rs.uniform(-1.5, 1.5, n), min_evals=10)


@pytest.mark.parametrize("seed", range(6))
def test_quadratic_plus_cosine(seed):
    rs = np.random.RandomState(100 + seed)
    n = rs.randint(2, 9)
    compare(quad(10.0 ** rs.uniform(-1, 3, n)), rs.uniform(-3, 3, n))


def gp_case(seed, m, kernel='rbf', mean_func=True, sig_fix=True):
    rs = np.random.RandomState(seed)
    x = np.sort(rs.uniform(0, 8, m))
    x[: max(1, m // 8)] = 0.0
    y = (48 - (4 + 3 * rs.rand()) * x + rs.randn(m) * 3 - 30) / 20
    return GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name=kernel, signal_variance=1.0,
                              signal_variance_fix=sig_fix, noise_variance=0.5, noise_variance_fix=False, mean_func=mean_func,
                              lengthscale=4.0, a_draw=abs(rs.randn()) if mean_func else None)


@pytest.mark.parametrize("seed,m,kernel,mean_func,sig_fix", [(0, 40, 'rbf', True, True), (1, 90, 'rbf', True, True),
                                                            (2, 150, 'rbf', False, True), (3, 60, 'rbf', True, False),
                                                            (4, 80, 'linear', True, True), (5, 200, 'rbf', True, True),
                                                            (6, 35, 'linear', False, False), (7, 300, 'rbf', True, True)])
def test_gp_refit_objective_same_evaluations_as_scipy(seed, m, kernel, mean_func, sig_fix, single_blas_thread):
    """The refit objective of a cluster (paramz _objective_grads: -(LML + log prior + log Jacobian) in Logexp space), from the
    cold start GPobs gives a new cluster (l = 4, noise_0 = 0.5, A = |randn|): the state machine asks for the same points as
    SciPy and stops at the same optimum."""
    o = gp_case(seed, m, kernel, mean_func, sig_fix)
    t0 = o.optimizer_array().copy()
    n = compare(lambda t: o.objective_grads(np.asarray(t, dtype=float)), t0, min_evals=5)
    print("evaluations:", n)


def test_warm_start_converges_at_once():
    """A refit from the previous optimum (the usual case inside the sampler): both stop after the same few evaluations."""
    o = gp_case(11, 120)
    t_opt, _f, _d = fmin_l_bfgs_b(lambda t: o.objective_grads(t), o.optimizer_array().copy(), maxfun=1000, maxiter=1000)
    fun = lambda t: o.objective_grads(np.asarray(t, dtype=float))   # noqa: E731
    n = compare(fun, t_opt, min_evals=1)
    assert n <= 6
