"""`MoGP_constrained` / `MoGP` — the reference's sampler API (mogp/mogp_constrained.py:27-347) on the CUDA engine.

Same constructor signature, `sample()` and `predict()`; the per-patient inner loop
(mogp_constrained.py:231-284) runs inside the engine's chain (`mogp_chain_step` / `mogp_chain_sweep`),
fed with the random draws the reference would consume from NumPy's global legacy generator, in the same
order (SURVEY.md §3.1 RNG ledger): per step one `randn(1,1)` (the newcomer's NegLinear map; none when
`mean_func=False`) and one `random_sample()` (inside `np.random.choice`).

Additions (all default to the reference's behaviour):
  refit        'move'  : L-BFGS-B refit after every move into an existing cluster (mogp_constrained.py:284)
               'sweep' : one refit per active cluster, in id order, at the end of each sweep
  init_z       initial partition instead of k-means (sklearn label numbering differs between versions)
  fast_updates None -> True for refit='sweep' (rank-n up/down-dates + leave-block-out own score),
               False for refit='move' (reference-exact data order; every move refactorises anyway)
  engine       an explicit `_capi.Engine` (one per chain / per GPU); by default every model creates its own
  optimizer    'native' (default: the engine's L-BFGS-B state machine, mogp_refit_clusters - the evaluation points of
               SciPy's fmin_l_bfgs_b, all clusters' refits in flight together) or 'scipy' (SciPy drives the objective)
  savepath     None (default) writes no checkpoints; the reference always saves (its default path is relative to its
               own source tree) - pass a directory to get its init / iter / MAP / final files
"""
import numbers
from pathlib import Path

import numpy as np
from scipy.special import logsumexp

from . import _capi
from .allocmodel import DPmix
from .obsmodel import GPobs, Standardize, make_spec, initial_theta, optimize_many, default_optimizer
from .utils import pack_patients


class MoGP_constrained:
    def __init__(self, X, Y, alpha, num_init_clusters=5, num_iter=100, rand_seed=0, savepath=None,
                 savename='MoGP_constrained', save_iter=25, kernel='rbf', signal_variance=1.,
                 signal_variance_fix=True, noise_variance=1., noise_variance_fix=False, mean_func=True,
                 threshold=None, onset_anchor=True, normalize=False, Y_mean=None, Y_std=None,
                 refit='move', init_z=None, fast_updates=None, engine=None, lookahead=0, optimizer=None):
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
        self.fast_updates = (refit == 'sweep') if fast_updates is None else bool(fast_updates)
        self._engine = engine
        if optimizer not in (None, 'native', 'scipy'):
            raise ValueError("optimizer must be 'native' or 'scipy'")
        self.optimizer = optimizer
        # refit='sweep' only: visits scored ahead per pass over the factors (0 = engine default, < 0 = off)
        self.lookahead = int(lookahead)

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
        self.trace = []
        # worker engines (streams + host threads) for the end-of-sweep refits: bounded by the host cores this
        # process can count on (several ranks share one box)
        import os
        cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 8)
        ranks_here = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1'))))
        # (the workers sleep in blocking syncs 80 % of the time, so a few more of them than cores per rank is fine:
        # 2 -> 4 workers is 1.54 -> 1.10 ms per evaluation, scripts/microbench.py)
        self.refit_concurrency = int(max(4, min(12, cores // ranks_here)))
        if os.environ.get('MOGP_REFIT_CONCURRENCY'):               # 1 = refit in the calling thread (profilers)
            self.refit_concurrency = max(1, int(os.environ['MOGP_REFIT_CONCURRENCY']))
        self._workers = None
        # native optimiser: clusters refitted together by mogp_refit_clusters (0 = the engine's default, 32 lanes); a process
        # that runs many chains at once gives each of them a few
        self.refit_lanes = int(os.environ.get('MOGP_REFIT_CONCURRENCY', '0'))
        self.min_margin = np.inf       # smallest |u - cdf edge| seen (assignment-parity diagnostic)
        self.n_refit_evals = 0

    # -- validation: mogp_constrained.py:92-144, same messages ---------------------------------------
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

    # -- normaliser: mogp_constrained.py:146-172 --------------------------------------------------------
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

    def __getstate__(self):
        st = dict(self.__dict__)
        st['_engine'] = None
        st['_workers'] = None
        return st

    # -- engine plumbing ---------------------------------------------------------------------------------
    @property
    def engine(self):
        # one private engine per model: an engine holds ONE chain, so two models trained in one process (the reference's
        # tutorial trains a constrained and an unconstrained one) must not share the process-wide default engine
        if self._engine is None:
            self._engine = _capi.Engine()
        return self._engine

    def _spec(self, mean_func=None):
        return make_spec(self.kernel, self.mean_func if mean_func is None else mean_func, self.signal_variance_fix,
                         self.noise_variance_fix)

    def _chain_config(self):
        cfg = _capi.ChainConfig()
        cfg.spec = self._spec()
        cfg.alpha = float(self.allocmodel.alpha)
        cfg.signal_variance = float(self.signal_variance)
        cfg.noise_variance = float(self.noise_variance)
        cfg.lengthscale = 4.0                                      # GPobs default (obsmodel.py:8)
        cfg.use_threshold = int(self.threshold is not None)
        cfg.threshold = float(self.threshold) if self.threshold is not None else 0.0
        cfg.thr_index = 1 if self.onset_anchor else 0              # mogp_constrained.py:256
        cfg.fast_updates = int(self.fast_updates)
        cfg.lookahead = int(getattr(self, 'lookahead', 0))
        return cfg

    def _view(self, slot):
        return GPobs(None, None, kernel=self.kernel, signal_variance_fix=self.signal_variance_fix,
                     noise_variance_fix=self.noise_variance_fix, mean_func=self.mean_func, engine=self.engine,
                     _slot=int(slot), optimizer=getattr(self, 'optimizer', None))

    def _sync_state(self):
        """Mirror z / Nk / obsmodel from the engine's chain."""
        z, Nk, slots = self.engine.chain_state(self.allocmodel.N)
        self.z = z
        self.allocmodel.Nk = Nk
        for k in range(len(Nk)):
            cur = self.obsmodel.get(k)
            if Nk[k] == 0 or slots[k] < 0:
                self.obsmodel[k] = None
            elif cur is None or cur.model.slot != slots[k]:
                self.obsmodel[k] = self._view(slots[k])
        self._slots = slots

    # -- chain snapshots (bench.py times different legs from the same state; not in the reference) ----------
    def snapshot(self):
        """Assignments, per-id hyper-parameters and the generator state of the running chain."""
        eng = self.engine
        z, Nk, slots = eng.chain_state(self.allocmodel.N)
        theta = np.zeros((len(Nk), 4))
        for k in range(len(Nk)):
            if Nk[k] > 0:
                theta[k] = eng.cluster_get_theta(int(slots[k]))
        return dict(z=z.copy(), theta=theta, rng=np.random.get_state())

    def restore(self, snap):
        """Re-attach the chain at a snapshot: every active cluster is refactorised from its members (in patient-index
        order) at the snapshot's hyper-parameters."""
        self.engine.chain_init(self._chain_config(), snap['z'], snap['theta'])
        np.random.set_state(snap['rng'])
        self.obsmodel = dict()
        self._sync_state()

    # -- initialisation: mogp_constrained.py:182-214 -----------------------------------------------------
    def initialize_sampler(self, init_K, onset_anchor):
        np.random.seed(self.rand_seed)
        if self.init_z is not None:
            z = np.asarray(self.init_z).astype(np.int32).copy()
        else:
            # KMeans(n_clusters=init_K, random_state=rand_seed, n_init=10) on each patient's first-visit time (:198-205),
            # on the device (mogp_kmeans_init: sklearn's arithmetic, draws and label numbering)
            col = 1 if onset_anchor else 0
            z = self.engine.kmeans_init(self.X[:, col], init_K, self.rand_seed, n_init=10).astype(np.int32)
        n_ids = len(np.unique(z))
        if z.max() + 1 != n_ids:
            raise ValueError("initial cluster ids must be 0..K-1 without gaps")
        # one randn per initial cluster, in id order (GPobs.__init__ -> NegLinear -> GPy.mappings.Linear)
        theta = np.zeros((n_ids, 4))
        for k in range(n_ids):
            a = np.random.randn(1, 1)[0, 0] if self.mean_func else None
            theta[k] = initial_theta(self.kernel, self.signal_variance, self.noise_variance, 4., a)
        # the NaN-padded matrices go to the engine as they are: compaction to the per-patient rows of
        # mogp_constrained.py:246-247 and the threshold cells Y[:, index_to_eval] (:256) are built on the device
        self.engine.set_patients_matrix(self.X, self.Y, 1 if onset_anchor else 0)
        self.engine.chain_init(self._chain_config(), z, theta)
        self.obsmodel = dict()
        self._sync_state()

    # -- the sampler: mogp_constrained.py:216-310 ------------------------------------------------------------
    def _draws_for_sweep(self, N):
        """permutation, then per step randn (if mean_func) and random_sample, interleaved as the reference
        consumes them (legacy randn caches its second polar variate, so the order matters)."""
        perm = np.random.permutation(np.arange(N))
        a = np.zeros(N)
        u = np.zeros(N)
        if self.mean_func:
            # standard_normal() is the draw behind randn(1, 1) (same legacy gauss stream, cached second variate included)
            gauss, unif = np.random.standard_normal, np.random.random_sample
            for s in range(N):
                a[s] = gauss()
                u[s] = unif()
        else:
            u[:] = np.random.random_sample(N)      # N consecutive doubles of the same stream
        return perm, a, u

    def sweep(self, perm=None, a=None, u=None):
        """One Gibbs sweep (mogp_constrained.py:231-284). Draws come from the global generator unless given."""
        N = self.allocmodel.N
        eng = self.engine
        if self.refit == 'move':
            if perm is None:
                perm = np.random.permutation(np.arange(N))
            for s, n in enumerate(perm):
                if a is not None:
                    an, un = a[s], u[s]
                else:
                    an = np.random.randn(1, 1)[0, 0] if self.mean_func else 0.0
                    un = np.random.random_sample()
                chosen, is_new, ids, p = eng.chain_step(n, an, un)
                self.p[n] = p
                self.min_margin = min(self.min_margin, eng.lib.mogp_chain_last_margin(eng.h))
                if not is_new:                                     # update_suffstats(optimize_params=True), :284
                    self._refit_id(chosen)
            self._sync_state()
        else:
            if perm is None:
                perm, a, u = self._draws_for_sweep(N)
            eng.chain_record_probs(True)
            _chosen, mm = eng.chain_sweep(perm, a if self.mean_func else None, u)
            self.min_margin = min(self.min_margin, mm)
            off, _ids, p = eng.chain_probs()                       # self.p[n] of every step (:276)
            for s, n in enumerate(perm):
                self.p[n] = p[off[s]:off[s + 1]]
            self._sync_state()
            self.refit_all()

    def refit_all(self):
        """refit='sweep': one L-BFGS-B refit per active cluster (independent; run concurrently on worker engines)."""
        models = [self.obsmodel[k].model for k in self._active()]
        native = (getattr(self, 'optimizer', None) or default_optimizer()) == 'native'
        optimize_many(models, [] if native else self._refit_workers(len(models)), 'native' if native else 'scipy',
                      max_concurrent=int(getattr(self, 'refit_lanes', 0)))
        for m in models:
            self.n_refit_evals += m.n_evals
            m.n_evals = 0

    def _refit_workers(self, n_models):
        want = min(self.refit_concurrency, n_models)
        if want <= 1:
            return []
        if not hasattr(self, '_workers') or self._workers is None:
            self._workers = []
        while len(self._workers) < want:
            self._workers.append(_capi.Engine(self.engine.device, blocking_sync=True))
        return self._workers[:want]

    def _refit_id(self, k):
        _z, _Nk, slots = self.engine.chain_state(self.allocmodel.N)
        view = self.obsmodel.get(k)
        if view is None or view.model.slot != slots[k]:
            view = self._view(slots[k])
            self.obsmodel[k] = view
        view.model.optimize()
        self.n_refit_evals += view.model.n_evals
        view.model.n_evals = 0

    def sample(self):
        self.initialize_sampler(self.num_init_clusters, self.onset_anchor)
        self.init_occupancy = self.allocmodel.Nk[self.allocmodel.Nk > 0].copy()
        if self.savepath is not None:
            Path.mkdir(self.savepath, parents=True, exist_ok=True)
        self._save_normalized_model("iterinit")
        for l in np.arange(1, self.num_iter):                      # num_iter - 1 sweeps (:228)
            self.sweep()
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

    fit = sample

    # -- prediction: mogp_constrained.py:312-347 -----------------------------------------------------------------
    def predict(self, x_new, y_new):
        """Membership probabilities of one new patient over every cluster id + the new-cluster slot."""
        self._check_validity_of_predict_inputs(x_new, y_new)
        if self.normalize:
            y_new = (y_new - self.mean) / self.std
            self._reset_normalizer()                               # :325-328
        # GPobs(x, y) with mean_func NOT forwarded -> default True -> one randn (:330-332); the optimize()
        # the reference runs next (:333) does not change the cached .ll it reads (:339) and draws nothing.
        a = np.random.randn(1, 1)[0, 0]
        p = self.predict_batch(np.asarray(x_new, dtype=np.float64).reshape(1, -1),
                               np.asarray(y_new, dtype=np.float64).reshape(1, -1), np.array([a]),
                               _normalized=True)[0]
        if self.normalize:
            self._apply_normalizer()                               # :343-345: the per-cluster normalisers end up attached
        return p

    PREDICT_BLOCK = 50000     # patients per block of the pipelined batch prediction

    def predict_batch(self, X_new, Y_new, a_draws, _normalized=False):
        """`predict` for P patients at once: NaN-padded P x T matrices -> P x (len(Nk)+1) probabilities.
        a_draws[p] = the randn the reference would draw for patient p's new-cluster model.
        Large batches go through in blocks: while the device scores block b (one host thread inside the engine, the GIL
        released), this thread compacts the rows of block b + 1 and turns the scores of block b - 1 into probabilities."""
        if self.normalize and not _normalized:
            Y_new = (Y_new - self.mean) / self.std
        eng = self.engine
        Nk = self.allocmodel.Nk
        active = self._active()
        slots = np.array([self.obsmodel[k].model.slot for k in active], dtype=np.int32)
        prior = np.log(np.hstack([Nk, self.allocmodel.alpha]) + 1e-16)                       # :335
        theta0 = initial_theta(self.kernel, self.signal_variance, self.noise_variance, 4., 0.0)
        spec_new = self._spec(mean_func=True)
        a_draws = np.asarray(a_draws, dtype=np.float64)

        def device(packed, a):
            off, x, y = packed
            return eng.score_pairs(off, x, y, slots), eng.new_cluster_ll(spec_new, theta0, off, x, y, a)   # :336-339

        def finish(res):
            pair, new = res
            score = np.tile(prior, (pair.shape[0], 1))
            score[:, active] += pair
            score[:, -1] += new
            if score.shape[0] == 1:
                return np.exp(score - logsumexp(score, axis=1, keepdims=True))               # :340-341
            # the same softmax for many rows without SciPy's per-call overhead: lse = max + log(sum(exp(s - max)))
            mx = np.max(score, axis=1, keepdims=True)
            lse = mx + np.log(np.sum(np.exp(score - mx), axis=1, keepdims=True))
            np.subtract(score, lse, out=score)
            return np.exp(score, out=score)

        P = X_new.shape[0]
        B = self.PREDICT_BLOCK
        if P <= B:
            return finish(device(pack_patients(X_new, Y_new), a_draws))
        from concurrent.futures import ThreadPoolExecutor
        out = np.empty((P, len(Nk) + 1))
        with ThreadPoolExecutor(max_workers=1) as gpu:
            pending = None                                         # (rows, future) of the block on the device
            for lo in range(0, P, B):
                hi = min(P, lo + B)
                packed = pack_patients(X_new[lo:hi], Y_new[lo:hi])
                fut = gpu.submit(device, packed, a_draws[lo:hi])
                if pending is not None:
                    out[pending[0]:pending[1]] = finish(pending[2].result())
                pending = (lo, hi, fut)
            out[pending[0]:pending[1]] = finish(pending[2].result())
        return out


class MoGP(MoGP_constrained):
    """'Unconstrained' MoGP = mean_func=False, threshold=None
    (example/tutorial_train_mogp_model.ipynb:L475-478)."""

    def __init__(self, X, Y, alpha, **kw):
        kw.setdefault('mean_func', False)
        kw.setdefault('threshold', None)
        super().__init__(X, Y, alpha, **kw)
