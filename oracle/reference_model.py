"""CPU restatement of the reference's sampler-side code — TEST INFRASTRUCTURE ONLY.

Follows, function by function:
  GPobs              /root/reference/mogp/obsmodel.py:6-94
  DPmix              /root/reference/mogp/allocmodel.py:5-34
  MoGP_constrained   /root/reference/mogp/mogp_constrained.py:27-347
  generate_toy_data, rank_cluster_prediction   /root/reference/mogp/utils.py:9-47
with GPy replaced by `oracle.gpy_restated`.  Random draws come from NumPy's
global legacy generator in exactly the reference's order (np.random.seed at
mogp_constrained.py:196; one randn(1,1) per GPobs with a mean function, via
GPy.mappings.Linear; one permutation per sweep at :231; one random_sample per
step inside np.random.choice at :277).

Two additions that the reference does not have, both off by default:
  * `refit`   : 'move' (reference: L-BFGS-B after every move into an existing
                cluster, mogp_constrained.py:284) or 'sweep' (all active clusters
                are refitted once, in id order, at the end of each sweep — the
                schedule BASELINE.json's 10k-patient configuration names).
  * `init_z`  : inject the initial partition instead of running k-means
                (sklearn's label numbering differs between versions; k-means is
                outside the hot path).
"""
import math
import numbers
from pathlib import Path

import numpy as np
from scipy.special import gammaln, logsumexp

from .gpy_restated import GPRegressionOracle, Standardize


# --------------------------------------------------------------------------
class GPobs:
    """One cluster's GP (obsmodel.py:6-94)."""

    def __init__(self, X, Y, kernel='rbf', signal_variance=1., signal_variance_fix=True, noise_variance=1.,
                 noise_variance_fix=False, mean_func=True, lengthscale=4.):
        # GPy.mappings.Linear.__init__ draws A from the global generator (neg_linear.py:19-20).
        a_draw = np.random.randn(1, 1)[0, 0] if mean_func else None
        self.model = GPRegressionOracle(X, Y, kernel_name=kernel, signal_variance=signal_variance,
                                        signal_variance_fix=signal_variance_fix, noise_variance=noise_variance,
                                        noise_variance_fix=noise_variance_fix, mean_func=mean_func,
                                        lengthscale=lengthscale, a_draw=a_draw)
        self.ll = self.model.log_likelihood()                     # obsmodel.py:56
        self.X, self.Y = X, Y
        self.tempX = self.tempY = self.templl = None

    def calc_ll(self):                                             # obsmodel.py:63-66
        self.ll = self.model.log_likelihood()
        return self.ll

    def update_data(self, X, Y):                                   # obsmodel.py:68-73
        self.X, self.Y = X, Y
        self.model.set_XY(X, Y)

    def calc_score(self, Xnew, Ynew):                              # obsmodel.py:75-81
        self.tempX = np.vstack([self.X, Xnew])
        self.tempY = np.vstack([self.Y, Ynew])
        return np.sum(self.model.log_predictive_density(Xnew, Ynew, None))

    def get_pred_at_loc(self, Xnew, index=0):                      # obsmodel.py:83-85
        return self.model.predict(Xnew[index].reshape(-1, 1))

    def update_suffstats(self, optimize_params=True):              # obsmodel.py:87-94
        self.X, self.Y = self.tempX, self.tempY
        self.model.set_XY(self.X, self.Y)
        if optimize_params:
            self.model.optimize()


# --------------------------------------------------------------------------
class DPmix:
    """CRP bookkeeping (allocmodel.py:5-34). Counts are int64 and bit-exact."""

    def __init__(self, alpha):
        self.alpha = alpha
        self.Nk = None
        self.N = None

    def calc_ll(self):                                             # allocmodel.py:12-16
        occ = self.Nk[self.Nk > 0]
        return (gammaln(self.alpha) - gammaln(self.N + self.alpha)
                + occ.shape[0] * np.log(self.alpha) + np.sum(gammaln(occ)))

    def calc_suffstats(self, z):                                   # allocmodel.py:18-20
        self.Nk = np.bincount(z)

    def calc_score(self, curr_k):                                  # allocmodel.py:22-27
        self.Nk[curr_k] -= 1
        weights = np.hstack([self.Nk, self.alpha])
        return np.log(weights + 1e-16), np.where(weights > 0)[0]

    def update_suffstats(self, new_k):                             # allocmodel.py:29-34
        if new_k > self.Nk.shape[0] - 1:
            self.Nk = np.hstack([self.Nk, 1])
        else:
            self.Nk[new_k] += 1


# --------------------------------------------------------------------------
def _compact(row):
    """Row of the NaN-padded matrix -> (n_i, 1) column of its non-NaN cells."""
    return row[~np.isnan(row)].reshape(-1, 1)


class MoGP_constrained:
    """Collapsed Gibbs sampler for the DP mixture of GPs (mogp_constrained.py:27-347)."""

    def __init__(self, X, Y, alpha, num_init_clusters=5, num_iter=100, rand_seed=0, savepath=None,
                 savename='MoGP_constrained', save_iter=25, kernel='rbf', signal_variance=1.,
                 signal_variance_fix=True, noise_variance=1., noise_variance_fix=False, mean_func=True,
                 threshold=None, onset_anchor=True, normalize=False, Y_mean=None, Y_std=None,
                 refit='move', init_z=None):
        self.X, self.Y = X, Y
        self.allocmodel = DPmix(alpha)
        self.num_iter, self.num_init_clusters, self.rand_seed = num_iter, num_init_clusters, rand_seed
        self.savepath = None if savepath is None else Path(savepath)
        self.savename, self.save_iter = savename, save_iter
        self.kernel, self.mean_func, self.threshold = kernel, mean_func, threshold
        self.signal_variance, self.signal_variance_fix = signal_variance, signal_variance_fix
        self.noise_variance, self.noise_variance_fix = noise_variance, noise_variance_fix
        self.onset_anchor, self.normalize = onset_anchor, normalize
        self.mean, self.std = Y_mean, Y_std
        if refit not in ('move', 'sweep'):
            raise ValueError("refit must be 'move' or 'sweep'")
        self.refit, self.init_z = refit, init_z

        self._check_validity_of_instance()
        self._initialize_normalizer()

        self.allocmodel.N, _ = X.shape
        self.z = np.zeros(self.allocmodel.N)
        self.p = dict()
        self.lalloc = np.zeros([self.num_iter])
        self.lobs = np.zeros([self.num_iter])
        self.obsmodel = dict()
        self.occupancy = {}
        self.best_ll = -np.inf
        self.ll = []
        self.burnin = 25
        self.trace = []            # per-sweep occupancies (what the reference logs at :286)
        self.step_log = None       # optional per-step record, see gibbs_step
        self.track_margin = False  # keep the smallest |u - cdf edge| of the chain in min_margin
        self.min_margin = np.inf

    # -- validation, mogp_constrained.py:92-144 (same messages) --------------------------
    def _check_validity_of_instance(self):
        a = self.allocmodel.alpha
        if not (isinstance(self.X, np.ndarray) and isinstance(self.Y, np.ndarray)):
            raise ValueError("X and Y must be np.ndarray. Wrong type passed in.")
        if self.X.shape != self.Y.shape:
            raise ValueError("X and Y must be the same size: {}, {}".format(self.X.shape, self.Y.shape))
        if not (isinstance(a, numbers.Number) and a >= 0):
            raise ValueError("parameter alpha must be a positive number. Was " + str(a))
        if not (isinstance(self.num_init_clusters, int) and self.num_init_clusters > 0):
            raise ValueError("number of initial clusters must be a positive integer greater than 0. Was "
                             + str(self.num_init_clusters))
        if not (isinstance(self.num_iter, int) and self.num_iter > 0):
            raise ValueError("number of iterations must be a positive integer greater than 0. Was "
                             + str(self.num_iter))
        if self.kernel not in ['rbf', 'linear']:
            raise ValueError("Implemented kernels limited to 'rbf' or 'linear'. Was " + str(self.kernel))
        for name in ('signal_variance', 'noise_variance'):
            if not isinstance(getattr(self, name), numbers.Number):
                raise ValueError("parameter {} must be number. Was {}".format(name, getattr(self, name)))
        for name in ('signal_variance_fix', 'noise_variance_fix', 'mean_func', 'onset_anchor', 'normalize'):
            if not isinstance(getattr(self, name), bool):
                raise ValueError("parameter {} must be boolean. Was {}".format(name, getattr(self, name)))
        if not (isinstance(self.threshold, numbers.Number) or self.threshold is None):
            raise ValueError("parameter threshold must be None or number. Was " + str(self.threshold))
        vals = self.Y[~np.isnan(self.Y)]
        if np.std(vals) == 0:
            raise ValueError('standard deviation of values is 0; check inputs')
        if self.normalize:
            if (self.mean is not None) and (self.std is None):
                raise ValueError("both mean and std must be passed to be used for norm. missing std.")
            if (self.mean is None) and (self.std is not None):
                raise ValueError("both mean and std must be passed to be used for norm. missing mean.")
        elif (self.mean is not None) and (self.std is not None):
            raise ValueError('if parameter normalize is false, mean and std cannot be provided')

    def _check_validity_of_predict_inputs(self, x_new, y_new):
        if not (isinstance(x_new, np.ndarray) and isinstance(self.Y, np.ndarray)):
            raise ValueError("x_new and y_new must be np.ndarray. Wrong type passed in.")
        if x_new.shape != y_new.shape:
            raise ValueError("x_new and y_new must be the same size: {} {}".format(x_new.shape, y_new.shape))

    # -- normaliser, mogp_constrained.py:146-172 ------------------------------------------
    def _initialize_normalizer(self):
        if self.normalize:
            if self.mean is None or self.std is None:
                vals = self.Y[~np.isnan(self.Y)]
                self.mean, self.std = np.mean(vals), np.std(vals)
            self.Y = (self.Y - self.mean) / self.std
            self.normalizer = Standardize(self.mean, self.std)
        else:
            self.normalizer = None

    def _active(self):
        return np.where(self.allocmodel.Nk > 0)[0]

    def _reset_normalizer(self):
        for k in self._active():
            self.obsmodel[k].model.normalizer = None

    def _apply_normalizer(self):
        for k in self._active():
            self.obsmodel[k].model.normalizer = self.normalizer

    def _save_normalized_model(self, file_suffix):                # mogp_constrained.py:174-179
        if self.savepath is None:
            return
        import joblib
        self._apply_normalizer()
        joblib.dump(self, self.savepath / "{}_{}.pkl".format(self.savename, file_suffix))
        self._reset_normalizer()

    # -- helpers ---------------------------------------------------------------------------------
    def _new_gpobs(self, Xc, Yc, mean_func=None):
        return GPobs(Xc, Yc, kernel=self.kernel, mean_func=self.mean_func if mean_func is None else mean_func,
                     signal_variance=self.signal_variance, signal_variance_fix=self.signal_variance_fix,
                     noise_variance=self.noise_variance, noise_variance_fix=self.noise_variance_fix)

    def _cluster_data(self, k):
        """Members' non-NaN cells in patient-index order, row-major (mogp_constrained.py:239-242)."""
        Xk, Yk = self.X[self.z == k], self.Y[self.z == k]
        return Xk[~np.isnan(Xk)].reshape(-1, 1), Yk[~np.isnan(Yk)].reshape(-1, 1)

    # -- initialisation, mogp_constrained.py:182-214 ----------------------------------------------
    def initialize_sampler(self, init_K, onset_anchor):
        np.random.seed(self.rand_seed)
        if self.init_z is not None:
            self.z = np.asarray(self.init_z).astype(np.int32).copy()
        else:
            from sklearn.cluster import KMeans
            col = 1 if onset_anchor else 0
            km = KMeans(n_clusters=init_K, random_state=self.rand_seed, n_init=10).fit(self.X[:, col].reshape(-1, 1))
            self.z = km.labels_
        for k in np.arange(len(np.unique(self.z))):
            self.obsmodel[k] = self._new_gpobs(*self._cluster_data(k))

    # -- one patient step, mogp_constrained.py:232-284 ----------------------------------------------
    def gibbs_step(self, n):
        curr_k = self.z[n]
        self.z[n] = -1
        if self.X[self.z == curr_k].shape[0] == 0:
            self.obsmodel[curr_k] = None
        else:
            self.obsmodel[curr_k].update_data(*self._cluster_data(curr_k))
        score, active = self.allocmodel.calc_score(curr_k)
        xn, yn = _compact(self.X[n]), _compact(self.Y[n])
        newcomer = self._new_gpobs(xn, yn)
        for k in active[:-1]:
            if self.threshold is not None:
                idx = 1 if self.onset_anchor else 0
                y_star, _ = self.obsmodel[k].get_pred_at_loc(xn, idx)
                if np.any((self.Y[n, idx] - y_star) > self.threshold):
                    score[k] = -np.inf
                    continue
            score[k] += self.obsmodel[k].calc_score(xn, yn)
        new_id = active[-1]
        score[new_id] += newcomer.ll
        s = score[active]
        self.p[n] = np.exp(s - logsumexp(s))
        margin = None
        if self.step_log is not None or self.track_margin:
            # the uniform np.random.choice is about to consume, peeked without disturbing the stream: distance of the
            # draw from the nearest cdf edge (how much score error the assignment tolerates)
            st = np.random.get_state()
            u = np.random.random_sample()
            np.random.set_state(st)
            cdf = np.cumsum(self.p[n])
            margin = float(np.min(np.abs(cdf / cdf[-1] - u)))
            self.min_margin = min(self.min_margin, margin)
        self.z[n] = np.random.choice(active, p=self.p[n])
        self.allocmodel.update_suffstats(self.z[n])
        if self.z[n] == new_id:
            self.obsmodel[self.z[n]] = newcomer
        else:
            self.obsmodel[self.z[n]].update_suffstats(optimize_params=(self.refit == 'move'))
        if self.step_log is not None:
            self.step_log.append(dict(n=int(n), curr_k=int(curr_k), active=active.copy(), score=s.copy(),
                                      p=self.p[n].copy(), z=int(self.z[n]), margin=margin))

    def end_of_sweep_refit(self):
        """refit='sweep' only: one L-BFGS-B refit per active cluster, in id order."""
        for k in self._active():
            self.obsmodel[k].model.optimize()

    # -- the sampler, mogp_constrained.py:216-310 --------------------------------------------------
    def sample(self):
        self.initialize_sampler(self.num_init_clusters, self.onset_anchor)
        self.allocmodel.calc_suffstats(self.z)
        idx = np.arange(self.allocmodel.N)
        self.init_occupancy = self.allocmodel.Nk[self.allocmodel.Nk > 0].copy()
        if self.savepath is not None:
            Path.mkdir(self.savepath, parents=True, exist_ok=True)
        self._save_normalized_model("iterinit")
        for l in np.arange(1, self.num_iter):                      # num_iter-1 sweeps (:228)
            for n in np.random.permutation(idx):
                self.gibbs_step(n)
            if self.refit == 'sweep':
                self.end_of_sweep_refit()
            self.trace.append(self.allocmodel.Nk[self.allocmodel.Nk > 0].copy())
            self.lalloc[l] = self.allocmodel.calc_ll()
            if l % self.save_iter == 0:
                self._save_normalized_model("iter{}".format(l))
            for k in self._active():
                self.lobs[l] += self.obsmodel[k].calc_ll()
            if l > self.burnin:
                self.occupancy[l] = self.allocmodel.Nk[self.allocmodel.Nk > 0]
                if self.lobs[l] + self.lalloc[l] > self.best_ll:
                    self.best_ll = self.lobs[l] + self.lalloc[l]
                    self._save_normalized_model("MAP")
        self.ll = self.lobs + self.lalloc
        self._apply_normalizer()
        if self.savepath is not None:
            import joblib
            joblib.dump(self, self.savepath / "{}.pkl".format(self.savename))

    # -- prediction, mogp_constrained.py:312-347 -------------------------------------------------
    def predict(self, x_new, y_new):
        self._check_validity_of_predict_inputs(x_new, y_new)
        if self.normalize:
            y_new = (y_new - self.mean) / self.std
            self._reset_normalizer()
        # mean_func is NOT forwarded here (:330-332) -> GPobs default True -> one randn draw.
        newcomer = self._new_gpobs(x_new.reshape(-1, 1), y_new.reshape(-1, 1), mean_func=True)
        ll_new = newcomer.ll            # cached before optimize(); the reference reads it stale (:339)
        newcomer.model.optimize()       # executed by the reference, result unused (:333)
        score = np.log(np.hstack([self.allocmodel.Nk, self.allocmodel.alpha]) + 1e-16)
        for k in self._active():
            score[k] += self.obsmodel[k].calc_score(x_new.reshape(-1, 1), y_new.reshape(-1, 1))
        score[-1] += ll_new
        p = np.exp(score - logsumexp(score))
        if self.normalize:
            self._apply_normalizer()
        return p


class MoGP(MoGP_constrained):
    """'Unconstrained' MoGP = mean_func=False, threshold=None
    (example/tutorial_train_mogp_model.ipynb:L475-478)."""

    def __init__(self, X, Y, alpha, **kw):
        kw.setdefault('mean_func', False)
        kw.setdefault('threshold', None)
        super().__init__(X, Y, alpha, **kw)


# --------------------------------------------------------------------------
def rank_cluster_prediction(model, x_new, y_new):                  # utils.py:9-29
    p = model.predict(x_new, y_new)
    ids = np.where(model.allocmodel.Nk > 0)[0]
    order = np.argsort(p[ids])[::-1]
    return ids[order], p[ids][order]


def generate_toy_data(seed=0):                                     # utils.py:32-47
    np.random.seed(seed=seed)
    blocks = [np.random.uniform(lo, hi, (5, 4)) for lo, hi in ((0.1, 7.), (3., 6.), (0.1, 3.), (0.1, 4.))]
    X_toy = np.sort(np.vstack(blocks))
    noise = np.random.randn(X_toy.shape[0], 1) * 4
    Ys = [48 / (1 + np.exp(-2 * (-X_toy + shift))) + noise for shift in (6, 3, 1)]
    return np.vstack([X_toy] * 3), np.vstack(Ys)
