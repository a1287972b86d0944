"""Weight-space form of the cluster GP (study for the next round).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a NumPy prototype that is checked against the dense restatement
(oracle/gpy_restated.py), so that the numbers quoted in DESIGN.md are reproducible.  Nothing in the product imports it.

The reference's model per cluster is GPRegression with an RBF kernel on ONE input dimension (mogp/obsmodel.py:24-34):
    Ky = k(X, X) + (noise + 1e-8) I,   k(x, x') = s2 exp(-(x - x')^2 / (2 l^2)).
On a bounded interval an analytic kernel is numerically low rank: interpolate it at r Chebyshev nodes t_1..t_r of an
interval [lo, hi] that contains the data,
    k(x, x') ~= c(x)^T K_tt c(x'),   c(x) = values at x of the r Lagrange polynomials of the nodes (barycentric form),
which is a FIXED feature map phi(x) = R c(x) with K_tt = R^T R (eigen-decomposition of the r x r node kernel, negative
rounding eigenvalues clipped).  Then with Phi = [phi(x_1) .. phi(x_m)]^T (m x r) and s = noise + 1e-8:
    Ky ~= Phi Phi^T + s I,           A = Phi^T Phi + s I_r  (r x r),
    mean(x*)  = m(x*) + phi*^T A^-1 Phi^T (y - m(X)),
    var(x*)   = k** - phi*^T phi* + s phi*^T A^-1 phi*          (Woodbury; NO cancellation: the first two terms differ by
                                                                 the truncation error only),
    log|Ky|   = log|A| + (m - r) log s,
    r^T Ky^-1 r = (r^T r - b^T A^-1 b) / s,  b = Phi^T r.
Everything a Gibbs step needs from a cluster is the r x r matrix A (a patient joining or leaving is a rank-n change of it)
and the r-vector b: O(r^2) per (visit, cluster) instead of O(m^2).
"""
import numpy as np

from .gpy_restated import GPY_JITTER, LOG_2_PI


def cheb_nodes(lo, hi, r):
    """Chebyshev points of the second kind on [lo, hi] and their barycentric weights."""
    j = np.arange(r)
    t = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * j / (r - 1))
    w = np.where(j % 2 == 0, 1.0, -1.0)
    w[0] *= 0.5
    w[-1] *= 0.5
    return t, w


def lagrange_values(x, t, w):
    """C[i, j] = l_j(x_i): the Lagrange basis of the nodes t at the points x (barycentric formula, exact at the nodes)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    d = x[:, None] - t[None, :]
    hit = d == 0.0
    d[hit] = 1.0
    q = w[None, :] / d
    C = q / np.sum(q, axis=1, keepdims=True)
    rows = np.any(hit, axis=1)
    C[rows] = hit[rows].astype(np.float64)
    return C


class RbfFeatures:
    """phi(x) for k(x, x') = variance exp(-(x - x')^2 / (2 lengthscale^2)) on [lo, hi], rank r."""

    def __init__(self, variance, lengthscale, lo, hi, r):
        self.variance, self.lengthscale = float(variance), float(lengthscale)
        self.t, self.w = cheb_nodes(lo, hi, r)
        Ktt = self.variance * np.exp(-0.5 * np.square(self.t[:, None] - self.t[None, :]) / self.lengthscale ** 2)
        lam, V = np.linalg.eigh(Ktt)
        lam = np.clip(lam, 0.0, np.inf)
        self.R = (V * np.sqrt(lam)[None, :]).T        # K_tt = R^T R

    def __call__(self, x):
        return lagrange_values(x, self.t, self.w) @ self.R.T   # (len(x), r)


class WeightSpaceGP:
    """The cluster GP in weight space: same quantities as GPRegressionOracle (rbf kernel, optional -A x mean)."""

    def __init__(self, X, Y, variance, lengthscale, noise, a_slope=None, r=48, lo=None, hi=None):
        x = np.asarray(X, dtype=np.float64).ravel()
        y = np.asarray(Y, dtype=np.float64).ravel()
        self.s = float(noise) + GPY_JITTER
        self.noise = float(noise)
        self.a = None if a_slope is None else float(a_slope)
        lo = float(np.min(x)) if lo is None else lo
        hi = float(np.max(x)) if hi is None else hi
        self.feat = RbfFeatures(variance, lengthscale, lo, hi, r)
        self.m = x.shape[0]
        self.Phi = self.feat(x)
        self.res = y - self._mean_fn(x)
        self.A = self.Phi.T @ self.Phi + self.s * np.eye(self.Phi.shape[1])
        self.LA = np.linalg.cholesky(self.A)
        self.b = self.Phi.T @ self.res
        self.Ainv_b = np.linalg.solve(self.LA.T, np.linalg.solve(self.LA, self.b))

    def _mean_fn(self, x):
        return 0.0 if self.a is None else -self.a * np.asarray(x, dtype=np.float64).ravel()   # neg_linear mapping (neg_linear.py)

    def log_likelihood(self):
        r = self.Phi.shape[1]
        logdet = 2.0 * np.sum(np.log(np.diag(self.LA))) + (self.m - r) * np.log(self.s)
        quad = (self.res @ self.res - self.b @ self.Ainv_b) / self.s
        return 0.5 * (-self.m * LOG_2_PI - logdet - quad)

    def predict(self, xs):
        xs = np.asarray(xs, dtype=np.float64).ravel()
        P = self.feat(xs)
        mu = self._mean_fn(xs) + P @ self.Ainv_b
        Z = np.linalg.solve(self.LA, P.T)                       # L_A^-1 phi*
        var = self.feat.variance - np.sum(P * P, axis=1) + self.s * np.sum(Z * Z, axis=0)
        var = np.clip(var, 1e-15, np.inf)                        # GPy clips the latent variance
        return mu, var + self.noise

    def log_predictive_density(self, xs, ys):
        mu, v = self.predict(xs)
        ys = np.asarray(ys, dtype=np.float64).ravel()
        return -0.5 * LOG_2_PI - 0.5 * np.log(v) - 0.5 * np.square(ys - mu) / v
