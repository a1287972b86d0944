"""`GPobs` — one cluster's GP, the drop-in boundary of the reference (mogp/obsmodel.py:6-94).

Same constructor, attributes (`model, ll, X, Y, tempX, tempY, templl`) and methods
(`calc_ll, update_data, calc_score, get_pred_at_loc, update_suffstats`); `.model` stands in for
the `GPy.models.GPRegression` object as far as the reference's callers use it
(`log_likelihood, set_XY, log_predictive_density, predict, optimize, normalizer, model[i]`).
Every number comes from a cluster slot of the CUDA engine (include/mogp_b200.h).
"""
import numpy as np
from scipy import optimize as _sciopt

from . import _capi


class Standardize:
    """GPy.util.normalizer.Standardize as used at mogp_constrained.py:155-158."""

    def __init__(self, mean=None, std=None):
        self.mean = mean
        self.std = std

    def inverse_mean(self, X):
        return X * self.std + self.mean

    def inverse_variance(self, var):
        return var * (self.std ** 2)


def _col(a):
    return np.asarray(a, dtype=np.float64).reshape(-1, 1)


class GPModel:
    """Device-backed stand-in for the GPRegression object held by `GPobs.model`.

    Parameter order = GPy's (mean function, kernel, likelihood), SURVEY.md §2.3:
        rbf    : [neg_linmap.A, rbf.variance, rbf.lengthscale, Gaussian_noise.variance]
        linear : [neg_linmap.A, sum.linear.variances, sum.bias.variance, Gaussian_noise.variance]
    without the leading A when mean_func=False (obsmodel.py:48-49).
    """

    def __init__(self, engine, spec, theta, slot=None, X=None, Y=None, optimizer=None):
        self._engine = engine
        self._spec = spec
        self.optimizer = optimizer     # 'native' | 'scipy' | None (= default_optimizer())
        self.normalizer = None
        self.n_evals = 0
        self._fail_count = 0
        self._last_grad = None
        self._slot = engine.cluster_create(spec, theta) if slot is None else slot
        self._owns_slot = slot is None
        if X is not None:
            self.set_XY(X, Y)

    # ---- structure -------------------------------------------------------------------------
    @property
    def slot(self):
        return self._slot

    @property
    def mean_func(self):
        return bool(self._spec.mean_func)

    @property
    def kernel_name(self):
        return 'rbf' if self._spec.kernel == _capi.KERNEL_RBF else 'linear'

    @property
    def param_names(self):
        k = ['rbf.variance', 'rbf.lengthscale'] if self.kernel_name == 'rbf' else \
            ['sum.linear.variances', 'sum.bias.variance']
        return (['neg_linmap.A'] if self.mean_func else []) + k + ['Gaussian_noise.variance']

    @property
    def param_array(self):
        th = self._engine.cluster_get_theta(self._slot)
        return th if self.mean_func else th[1:]

    def __getitem__(self, i):
        return self.param_array[i]

    @property
    def X(self):
        return self._engine.cluster_get_data(self._slot)[0].reshape(-1, 1)

    @property
    def Y(self):
        return self._engine.cluster_get_data(self._slot)[1].reshape(-1, 1)

    # ---- the five GPy calls of obsmodel.py ---------------------------------------------------
    def set_XY(self, X, Y):                                       # obsmodel.py:73,92
        self._engine.cluster_set_data(self._slot, X, Y)

    def log_likelihood(self):                                     # obsmodel.py:56,65
        return self._engine.cluster_loglik(self._slot)

    def log_predictive_density(self, x_test, y_test, Y_metadata=None):   # obsmodel.py:80
        return self._engine.cluster_lpd(self._slot, x_test, y_test).reshape(-1, 1)

    def predict(self, Xnew):                                      # obsmodel.py:85
        mu, var = self._engine.cluster_predict(self._slot, Xnew)
        mu, var = mu.reshape(-1, 1), var.reshape(-1, 1)
        if self.normalizer is not None:
            mu = self.normalizer.inverse_mean(mu)
            var = self.normalizer.inverse_variance(var)
        return mu, var

    def objective_function(self):
        return self._engine.cluster_objective(self._slot, None, want_grad=False)[0]

    @property
    def n_free(self):
        """Free hyper-parameters: absent A (mean_func=False) and fixed variances are excluded (obsmodel.py:29-30, 51-52)."""
        s = self._spec
        return int(bool(s.mean_func)) + int(not s.signal_variance_fix) + 1 + int(not s.noise_variance_fix)

    def objective_grads(self, t, worker=None):
        """paramz Model._objective_grads: (f, df/dt) in Logexp space; a non-PD covariance returns a huge
        objective up to 10 times in a row, as paramz does."""
        self.n_evals += 1
        try:
            f, g = self._engine.cluster_objective(self._slot, t, worker=worker, nfree=self.n_free)
            self._fail_count = 0
            self._last_grad = g
        except np.linalg.LinAlgError:
            if self._fail_count >= 10:
                raise
            self._fail_count += 1
            f = np.finfo(np.float64).max
            g = np.clip(self._last_grad if self._last_grad is not None else np.zeros(len(t)), -1e10, 1e10)
        return f, g

    def optimize(self, worker=None, t0=None):                     # obsmodel.py:94, mogp_constrained.py:333
        """paramz Model.optimize() default = L-BFGS-B (SciPy's fmin_l_bfgs_b: maxfun = maxiter = 1000, m=10, factr=1e7,
        pgtol=1e-5), warm-started at the current hyper-parameters; then f(x_opt) and a final set.
        optimizer 'native' (default): the engine's own L-BFGS-B state machine behind mogp_refit_clusters - the same
        evaluation points as SciPy's (tests/test_lbfgsb.py), no Python between evaluations; 'scipy': SciPy drives
        mogp_cluster_objective (`worker`: on another engine's stream, see `optimize_many`; `t0`: the starting point read by
        the caller's thread)."""
        if self.n_free == 0:
            return
        if (self.optimizer or default_optimizer()) == 'native' and worker is None:
            self.n_evals += int(self._engine.refit_clusters([self._slot])[0])
            return
        if t0 is None:
            t0 = self._engine.cluster_optimizer_array(self._slot)
        x_opt, _f, _d = _sciopt.fmin_l_bfgs_b(lambda t: self.objective_grads(t, worker), t0, maxfun=1000, maxiter=1000)
        self.objective_grads(x_opt, worker)
        self._engine.cluster_objective(self._slot, x_opt, want_grad=False, worker=worker, nfree=self.n_free)

    def plot_mean(self, *a, **k):
        raise NotImplementedError("GPy/matplotlib plotting is outside the likelihood engine")

    plot_confidence = plot_mean

    def release(self):
        if self._owns_slot and self._slot is not None:
            try:
                self._engine.cluster_free(self._slot)
            except Exception:
                pass
            self._slot = None

    def __del__(self):
        self.release()

    # ---- pickling (joblib.dump of the whole sampler, mogp_constrained.py:174-179) ----------------
    def __getstate__(self):
        x, y = self._engine.cluster_get_data(self._slot)
        s = self._spec
        return dict(theta=self._engine.cluster_get_theta(self._slot), x=x, y=y, normalizer=self.normalizer,
                    spec=(s.kernel, s.mean_func, s.signal_variance_fix, s.noise_variance_fix))

    def __setstate__(self, st):
        self._engine = _capi.default_engine()
        self._spec = _capi.ModelSpec(*st['spec'])
        self.normalizer = st['normalizer']
        self.n_evals, self._fail_count, self._last_grad = 0, 0, None
        self.optimizer = None
        self._slot = self._engine.cluster_create(self._spec, st['theta'])
        self._owns_slot = True
        self.set_XY(st['x'], st['y'])


def default_optimizer():
    """'native' (the engine's L-BFGS-B, mogp_refit_clusters) unless MOGP_OPTIMIZER=scipy."""
    import os
    v = os.environ.get('MOGP_OPTIMIZER', 'native')
    if v not in ('native', 'scipy'):
        raise ValueError("MOGP_OPTIMIZER must be 'native' or 'scipy'")
    return v


def optimize_many(models, workers, optimizer=None, max_concurrent=0):
    """Refit several clusters of one engine concurrently: the L-BFGS-B runs are independent (the reference
    refits one cluster at a time, obsmodel.py:94; the result per cluster is the same).
    'native' (default): one call of mogp_refit_clusters - a lane (stream + workspace) and an L-BFGS-B state machine per
    cluster, all driven by this thread.  'scipy' (parity mode): each cluster runs SciPy's optimiser in its own host thread
    on a worker engine's stream.  Either way every cluster sees exactly the evaluations of its own sequential `optimize()`."""
    models = list(models)
    if not models:
        return
    if (optimizer or default_optimizer()) == 'native':
        live = [m for m in models if m.n_free > 0]
        if live:
            n = live[0]._engine.refit_clusters([m.slot for m in live], max_concurrent)
            for m, k in zip(live, n):
                m.n_evals += int(k)
        return
    if not workers or len(models) == 1:
        for m in models:
            m.optimize()
        return
    import queue
    from concurrent.futures import ThreadPoolExecutor
    models[0]._engine.sync()                       # the owner's stream must be idle before workers read its clusters
    free = queue.Queue()
    for w in workers:
        free.put(w)

    # the owner engine is only touched from this thread: the starting points are read here, before the dispatch
    starts = {id(m): m._engine.cluster_optimizer_array(m.slot) for m in models if m.n_free > 0}

    def run(m):
        w = free.get()
        try:
            if m.n_free > 0:
                m.optimize(worker=w, t0=starts[id(m)])
        finally:
            free.put(w)

    with ThreadPoolExecutor(max_workers=len(workers)) as pool:
        list(pool.map(run, models))


def make_spec(kernel, mean_func, signal_variance_fix, noise_variance_fix):
    if kernel not in _capi.KERNEL_CODE:
        raise ValueError("Implemented kernels limited to 'rbf' or 'linear'. Was " + str(kernel))
    return _capi.ModelSpec(_capi.KERNEL_CODE[kernel], int(bool(mean_func)), int(bool(signal_variance_fix)),
                           int(bool(noise_variance_fix)))


def initial_theta(kernel, signal_variance, noise_variance, lengthscale, a_draw):
    """Starting hyper-parameters of a fresh GPobs (obsmodel.py:25-54): A0 = |randn| (GPy.mappings.Linear draw
    mapped to the positive side by the Logexp constraint), l0 = lengthscale (4), Bias variance 1."""
    k1 = float(lengthscale) if kernel == 'rbf' else 1.0
    return np.array([abs(float(a_draw)) if a_draw is not None else 0.0, float(signal_variance), k1,
                     float(noise_variance)])


class GPobs:
    """Reference signature: obsmodel.py:7-8."""

    def __init__(self, X, Y, kernel='rbf', signal_variance=1., signal_variance_fix=True, noise_variance=1.,
                 noise_variance_fix=False, mean_func=True, lengthscale=4., engine=None, _slot=None, optimizer=None):
        engine = engine or _capi.default_engine()
        spec = make_spec(kernel, mean_func, signal_variance_fix, noise_variance_fix)
        if _slot is not None:               # view of a slot the chain engine already manages
            self.model = GPModel(engine, spec, None, slot=_slot, optimizer=optimizer)
            self._X, self._Y = None, None
        else:
            # GPy.mappings.Linear.__init__ draws A from the global generator (neg_linear.py:19-20)
            a_draw = np.random.randn(1, 1)[0, 0] if mean_func else None
            theta = initial_theta(kernel, signal_variance, noise_variance, lengthscale, a_draw)
            self.model = GPModel(engine, spec, theta, X=_col(X), Y=_col(Y), optimizer=optimizer)
            self._X, self._Y = X, Y
        self.ll = self.model.log_likelihood()                      # obsmodel.py:56
        self.tempX = None
        self.tempY = None
        self.templl = None

    # X / Y: the arrays the caller handed over, or - for a view of a chain-managed slot - the engine's copy
    @property
    def X(self):
        return self.model.X if self._X is None else self._X

    @X.setter
    def X(self, v):
        self._X = v

    @property
    def Y(self):
        return self.model.Y if self._Y is None else self._Y

    @Y.setter
    def Y(self, v):
        self._Y = v

    def calc_ll(self):                                             # obsmodel.py:63-66
        self.ll = self.model.log_likelihood()
        return self.ll

    def update_data(self, X, Y):                                   # obsmodel.py:68-73
        self.X, self.Y = X, Y
        self.model.set_XY(_col(X), _col(Y))

    def calc_score(self, Xnew, Ynew):                              # obsmodel.py:75-81
        X, Y = _col(self.X), _col(self.Y)
        self.tempX = np.vstack([X, _col(Xnew)])
        self.tempY = np.vstack([Y, _col(Ynew)])
        return np.sum(self.model.log_predictive_density(_col(Xnew), _col(Ynew), None))

    def get_pred_at_loc(self, Xnew, index=0):                      # obsmodel.py:83-85
        return self.model.predict(np.asarray(Xnew)[index].reshape(-1, 1))

    def update_suffstats(self, optimize_params=True):              # obsmodel.py:87-94
        self.X, self.Y = self.tempX, self.tempY
        self.model.set_XY(self.X, self.Y)
        if optimize_params:
            self.model.optimize()
