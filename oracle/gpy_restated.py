"""Restatement of the GPy 1.10 / paramz 0.9.5 arithmetic reached by the reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference
(`/root/reference/mogp/obsmodel.py`) builds one `GPy.models.GPRegression` per
cluster and calls five of its methods: `log_likelihood` (obsmodel.py:56,65),
`set_XY` (:73,92), `log_predictive_density` (:80), `predict` (:85) and
`optimize` (:94).  GPy is a third-party dependency that is NOT under
/root/reference (setup.py:14 pins GPy==1.10) and is not installable here, so
its published algorithm is restated below with NumPy + the same LAPACK
routines GPy reaches through SciPy.  Each block names the GPy routine whose
behaviour it follows; the result is pinned end-to-end by the reference's
executed tutorial notebook (tests/test_oracle_kat.py).

Parameter order of the flat parameter vector (GPy: mean_function, kernel,
likelihood; confirmed by analysis/analysis_utils.py:386 and the notebook
table, example/tutorial_train_mogp_model.ipynb:L286-289):
    rbf    : [neg_linmap.A, rbf.variance, rbf.lengthscale, Gaussian_noise.variance]
    linear : [neg_linmap.A, sum.linear.variances, sum.bias.variance, Gaussian_noise.variance]
(the leading A is absent when mean_func=False, obsmodel.py:48-49).
"""
import math

import numpy as np
from scipy import optimize as _sciopt
from scipy.linalg import lapack as _lapack
from scipy.special import gammaln as _gammaln

LOG_2_PI = math.log(2.0 * math.pi)
_LIM_VAL = 36.0                                  # paramz.transformations._lim_val
_LOG_LIM_VAL = math.log(np.finfo(np.float64).max)  # paramz.transformations._log_lim_val
GPY_JITTER = 1e-8                                # ExactGaussianInference: diag.add(Ky, variance + 1e-8)


# --------------------------------------------------------------------------
# paramz.transformations.Logexp
# --------------------------------------------------------------------------
def logexp_f(t):
    """theta = log(1 + exp(t)); identity above _lim_val (paramz Logexp.f)."""
    t = np.asarray(t, dtype=np.float64)
    return np.where(t > _LIM_VAL, t, np.log1p(np.exp(np.clip(t, -_LOG_LIM_VAL, _LIM_VAL))))


def logexp_finv(theta):
    """t = log(expm1(theta)); identity above _lim_val (paramz Logexp.finv)."""
    theta = np.asarray(theta, dtype=np.float64)
    with np.errstate(over='ignore'):
        return np.where(theta > _LIM_VAL, theta, np.log(np.expm1(theta)))


def logexp_gradfactor(theta):
    """d theta / d t = -expm1(-theta) (paramz Logexp.gradfactor)."""
    theta = np.asarray(theta, dtype=np.float64)
    return np.where(theta > _LIM_VAL, 1.0, -np.expm1(-theta))


def logexp_log_jacobian(theta):
    """ln|d theta/d t| as paramz writes it: ln(expm1(theta)) - theta."""
    return logexp_finv(theta) - theta


def logexp_log_jacobian_grad(theta):
    return 1.0 / np.expm1(theta)


# --------------------------------------------------------------------------
# GPy.priors.Gamma
# --------------------------------------------------------------------------
class GammaPrior:
    """GPy.priors.Gamma: lnpdf(x) = -lnGamma(a) + a ln b + (a-1) ln x - b x."""

    def __init__(self, a, b):
        self.a = float(a)
        self.b = float(b)
        self.constant = -_gammaln(self.a) + self.a * math.log(self.b)

    @classmethod
    def from_EV(cls, E, V):
        """GPy.priors.Gamma.from_EV: a = E^2/V, b = E/V (obsmodel.py:28,32,47,54)."""
        return cls(float(E) ** 2 / float(V), float(E) / float(V))

    def lnpdf(self, x):
        return self.constant + (self.a - 1.0) * math.log(x) - self.b * x

    def lnpdf_grad(self, x):
        return (self.a - 1.0) / x - self.b


# --------------------------------------------------------------------------
# GPy.util.normalizer.Standardize (mogp_constrained.py:155-158)
# --------------------------------------------------------------------------
class Standardize:
    def __init__(self, mean=None, std=None):
        self.mean = mean
        self.std = std

    def inverse_mean(self, X):
        return X * self.std + self.mean

    def inverse_variance(self, var):
        return var * (self.std ** 2)


# --------------------------------------------------------------------------
# GPy.util.linalg
# --------------------------------------------------------------------------
def jitchol(A, maxtries=5):
    """GPy.util.linalg.jitchol: dpotrf, then up to `maxtries` jittered retries."""
    A = np.ascontiguousarray(A)
    L, info = _lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            return np.linalg.cholesky(A + np.eye(A.shape[0]) * jitter)
        except np.linalg.LinAlgError:
            jitter *= 10.0
        finally:
            num_tries += 1
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def pdinv(A):
    """GPy.util.linalg.pdinv: (Ai, L, Li, logdet) via dpotrf / dtrtri / dpotri."""
    L = jitchol(A)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    Li = _lapack.dtrtri(L, lower=1)[0]
    Ai, _ = _lapack.dpotri(L, lower=1)
    # symmetrify: copy the lower triangle to the upper one
    il = np.tril_indices(Ai.shape[0], -1)
    Ai.T[il] = Ai[il]
    return Ai, L, Li, logdet


# --------------------------------------------------------------------------
# kernels: GPy.kern.RBF, GPy.kern.Linear + GPy.kern.Bias
# --------------------------------------------------------------------------
def _unscaled_dist(X, X2=None):
    """GPy.kern.src.stationary.Stationary._unscaled_dist for 1-D inputs.

    r^2 = x^2 + x'^2 - 2 x x' by the EXPANSION (not (x-x')^2), diagonal forced
    to zero in the symmetric case, clipped at 0, then sqrt.
    """
    x = X[:, 0]
    if X2 is None:
        xsq = np.square(x)
        r2 = -2.0 * np.outer(x, x) + (xsq[:, None] + xsq[None, :])
        r2[np.diag_indices(x.shape[0])] = 0.0
        r2 = np.clip(r2, 0.0, np.inf)
        return np.sqrt(r2)
    x2 = X2[:, 0]
    r2 = -2.0 * np.outer(x, x2) + (np.square(x)[:, None] + np.square(x2)[None, :])
    r2 = np.clip(r2, 0.0, np.inf)
    return np.sqrt(r2)


class GPRegressionOracle:
    """GPy.models.GPRegression as configured by `GPobs.__init__` (obsmodel.py:25-54).

    kernel_name 'rbf'   : RBF(variance=signal_variance, lengthscale) (obsmodel.py:27)
    kernel_name 'linear': Linear(variances=signal_variance) + Bias(1) (obsmodel.py:37)
    mean_func           : NegLinear, m(x) = -A x, A0 = |a_draw| (neg_linear.py:19-23;
                          GPy.mappings.Linear draws A ~ randn(1,1); the Gamma prior's
                          Logexp constraint maps a negative start to abs()).
    """

    def __init__(self, X, Y, kernel_name='rbf', signal_variance=1.0, signal_variance_fix=True,
                 noise_variance=1.0, noise_variance_fix=False, mean_func=True, lengthscale=4.0,
                 a_draw=None):
        self.kernel_name = kernel_name
        self.mean_func = bool(mean_func)
        names, values, fixed, priors = [], [], [], []
        if self.mean_func:
            if a_draw is None:
                raise ValueError("mean_func=True needs the randn(1,1) draw of GPy.mappings.Linear")
            names.append('neg_linmap.A')
            values.append(abs(float(a_draw)))
            fixed.append(False)
            priors.append(GammaPrior.from_EV(2.0 / 3.0, 0.2))            # obsmodel.py:47
        if kernel_name == 'rbf':
            names += ['rbf.variance', 'rbf.lengthscale']
            values += [float(signal_variance), float(lengthscale)]
            fixed += [bool(signal_variance_fix), False]
            priors += [None if signal_variance_fix else GammaPrior.from_EV(1.0, 0.5),   # obsmodel.py:29-32
                       GammaPrior.from_EV(4.0, 9.0)]                                      # obsmodel.py:28
        elif kernel_name == 'linear':
            names += ['sum.linear.variances', 'sum.bias.variance']
            values += [float(signal_variance), 1.0]
            fixed += [bool(signal_variance_fix), False]
            priors += [None if signal_variance_fix else GammaPrior.from_EV(1.0, 0.5),   # obsmodel.py:38-41
                       None]                                                              # Bias: +ve, no prior
        else:
            raise ValueError("kernel must be 'rbf' or 'linear'")
        names.append('Gaussian_noise.variance')
        values.append(float(noise_variance))
        fixed.append(bool(noise_variance_fix))
        priors.append(None if noise_variance_fix else GammaPrior.from_EV(0.75, 0.25 ** 2))  # obsmodel.py:51-54
        self.param_names = names
        self.param_array = np.array(values, dtype=np.float64)
        self.fixed = np.array(fixed, dtype=bool)
        self.priors = priors
        self.normalizer = None
        self.n_evals = 0          # objective evaluations (bookkeeping for the CPU baseline)
        self._fail_count = 0
        self.set_XY(X, Y)

    # ---- parameter access ---------------------------------------------------
    def __getitem__(self, i):
        return self.param_array[i]

    @property
    def _off(self):
        return 1 if self.mean_func else 0

    @property
    def A(self):
        return self.param_array[0] if self.mean_func else 0.0

    @property
    def noise(self):
        return self.param_array[-1]

    # ---- kernel ---------------------------------------------------------------
    def _K(self, X, X2=None):
        k0, k1 = self.param_array[self._off], self.param_array[self._off + 1]
        if self.kernel_name == 'rbf':
            r = _unscaled_dist(X, X2) / k1
            return k0 * np.exp(-0.5 * r ** 2), r
        x = X[:, 0]
        x2 = x if X2 is None else X2[:, 0]
        xx = np.outer(x, x2)
        return k0 * xx + k1, xx

    def _Kdiag(self, X):
        k0, k1 = self.param_array[self._off], self.param_array[self._off + 1]
        if self.kernel_name == 'rbf':
            return np.full(X.shape[0], k0)
        return k0 * np.square(X[:, 0]) + k1

    def _mean(self, X):
        if self.mean_func:
            return np.dot(X, -np.array([[self.A]]))                      # neg_linear.py:22-23
        return 0.0

    # ---- GP.set_XY / parameters_changed -------------------------------------------
    def set_XY(self, X, Y):
        self.X = np.asarray(X, dtype=np.float64).reshape(-1, 1)
        self.Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
        self._inference()

    def _inference(self):
        """ExactGaussianInference.inference + the gradient updates of GP.parameters_changed."""
        X, Y = self.X, self.Y
        n = Y.shape[0]
        resid = Y - self._mean(X)
        K, aux = self._K(X)
        Ky = K.copy()
        Ky[np.diag_indices(n)] += self.noise + GPY_JITTER
        Wi, L, _Li, logdet = pdinv(Ky)
        alpha, _ = _lapack.dpotrs(L, resid, lower=1)
        self._L = L
        self._alpha = alpha
        self._lml = 0.5 * (-n * LOG_2_PI - logdet - np.sum(alpha * resid))
        dL_dK = 0.5 * (np.dot(alpha, alpha.T) - Wi)
        g = np.zeros_like(self.param_array)
        o = self._off
        if self.mean_func:
            g[0] = np.dot(-X.T, alpha)[0, 0]                                  # neg_linear.py:25-26
        if self.kernel_name == 'rbf':
            g[o] = np.sum(K * dL_dK) / self.param_array[o]                    # Stationary: variance
            g[o + 1] = np.sum(dL_dK * K * aux * aux) / self.param_array[o + 1]  # -sum(dL_dr*r)/l, dK_dr=-rK
        else:
            g[o] = np.sum(aux * dL_dK)                                        # Linear.variances
            g[o + 1] = np.sum(dL_dK)                                          # Bias.variance
        g[-1] = np.trace(dL_dK)                                               # Gaussian.exact_inference_gradients
        self._grad_lml = g

    def log_likelihood(self):
        """GP.log_likelihood: the marginal likelihood, WITHOUT the prior term."""
        return float(self._lml)

    # ---- prediction -------------------------------------------------------------------
    def _raw_predict(self, Xnew):
        """Posterior._raw_predict (diagonal variance) + mean function."""
        Xnew = np.asarray(Xnew, dtype=np.float64).reshape(-1, 1)
        Kx, _ = self._K(self.X, Xnew)
        mu = np.dot(Kx.T, self._alpha)
        tmp = _lapack.dtrtrs(self._L, Kx, lower=1)[0]
        var = (self._Kdiag(Xnew) - np.square(tmp).sum(0))[:, None]
        var = np.clip(var, 1e-15, np.inf)
        if self.mean_func:
            mu = mu + self._mean(Xnew)
        return mu, var

    def log_predictive_density(self, x_test, y_test, Y_metadata=None):
        """GP.log_predictive_density -> Gaussian.log_predictive_density, one value per visit."""
        mu, var = self._raw_predict(x_test)
        v = var + self.noise
        y_test = np.asarray(y_test, dtype=np.float64).reshape(-1, 1)
        return -0.5 * LOG_2_PI - 0.5 * np.log(v) - 0.5 * np.square(y_test - mu) / v

    def predict(self, Xnew):
        """GP.predict: (mean, var + noise), un-normalised when a normalizer is attached."""
        mu, var = self._raw_predict(Xnew)
        var = var + self.noise
        if self.normalizer is not None:
            mu = self.normalizer.inverse_mean(mu)
            var = self.normalizer.inverse_variance(var)
        return mu, var

    # ---- objective in optimiser space (paramz Model._objective_grads) ---------------------
    def _free(self):
        return np.where(~self.fixed)[0]

    def log_prior(self):
        lp = 0.0
        for i, pr in enumerate(self.priors):
            if pr is not None:
                th = self.param_array[i]
                lp += pr.lnpdf(th) + float(logexp_log_jacobian(th))
        return lp

    def _log_prior_grad(self):
        g = np.zeros_like(self.param_array)
        for i, pr in enumerate(self.priors):
            if pr is not None:
                th = self.param_array[i]
                g[i] = pr.lnpdf_grad(th) + float(logexp_log_jacobian_grad(th))
        return g

    def objective(self):
        return -self._lml - self.log_prior()

    def optimizer_array(self):
        return logexp_finv(self.param_array[self._free()])

    def _set_optimizer_array(self, t):
        self.param_array[self._free()] = logexp_f(t)
        self._inference()

    def objective_grads(self, t):
        """(f, df/dt) at optimiser coordinates t; failures handled as paramz does."""
        self.n_evals += 1
        try:
            self._set_optimizer_array(np.asarray(t, dtype=np.float64))
            f = self.objective()
            free = self._free()
            g = -(self._grad_lml + self._log_prior_grad())[free] * logexp_gradfactor(self.param_array[free])
            self._fail_count = 0
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= 10:
                raise
            self._fail_count += 1
            f = np.finfo(np.float64).max
            free = self._free()
            g = np.clip(-(self._grad_lml + self._log_prior_grad())[free]
                        * logexp_gradfactor(self.param_array[free]), -1e10, 1e10)
        return float(f), np.asarray(g, dtype=np.float64)

    def optimize(self):
        """paramz Model.optimize() default: scipy L-BFGS-B, maxfun = maxiter = 1000.

        paramz.optimization.opt_lbfgsb: fmin_l_bfgs_b(f_fp, x0, maxfun=1000, maxiter=1000)
        with SciPy's defaults m=10, factr=1e7, pgtol=1e-5; then f_opt = f_fp(x_opt)[0] and
        `optimizer_array = x_opt` (one more inference).
        """
        if self._free().size == 0:
            return
        t0 = self.optimizer_array()
        x_opt, _f, _d = _sciopt.fmin_l_bfgs_b(self.objective_grads, t0, maxfun=1000, maxiter=1000)
        self.objective_grads(x_opt)
        self._set_optimizer_array(x_opt)
