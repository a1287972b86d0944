/*
 * mogp_b200.h — C ABI of the B200 (sm_100a) likelihood engine for MoGP.
 *
 * The reference (fraenkel-lab/mogp) has no FFI: its hot path is a Python duck-typed
 * interface, `GPobs` (mogp/obsmodel.py:6-94), called by the collapsed Gibbs sampler
 * (mogp/mogp_constrained.py:231-284) and by `predict` (:312-347), with all arithmetic
 * inside GPy.  Each entry point below names the reference call it replaces.  A
 * maintainer binds these with ctypes; `mogp_b200/_capi.py` is
 * that binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MOGP_E* code; the message is
 *     available from mogp_last_error(); nothing throws across the ABI;
 *   - all `const double*` / `const int64_t*` arguments are HOST pointers unless the
 *     name ends in `_dev`; copies to / from the device happen inside the call on the
 *     engine's stream (so timing a call with host buffers is an end-to-end timing);
 *   - an engine owns its CUDA streams (a main one, and a side one on which a Gibbs step runs its
 *     independent pieces); calls on one engine are not thread-safe, calls on different engines
 *     are (one host thread + one engine per Gibbs chain; worker engines for concurrent refits,
 *     see mogp_cluster_objective_on);
 *   - theta is the natural-parameter vector in GPy's order
 *         rbf    : [A, rbf.variance, rbf.lengthscale, Gaussian_noise.variance]
 *         linear : [A, linear.variances, bias.variance, Gaussian_noise.variance]
 *     (A is present but ignored when mean_func = 0);
 *   - all arithmetic is FP64.
 */
#ifndef MOGP_B200_H
#define MOGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOGP_OK 0
#define MOGP_EINVAL (-1)  /* bad argument */
#define MOGP_ECUDA (-2)   /* CUDA runtime error (no device, launch failure, ...) */
#define MOGP_ENOMEM (-3)  /* device allocation failed */
#define MOGP_ENOTPD (-4)  /* covariance not positive definite (GPy raises LinAlgError) */
#define MOGP_EPROB (-5)   /* probabilities contain NaN (np.random.choice raises ValueError) */
#define MOGP_ESTATE (-6)  /* call sequence error (e.g. cluster has no data) */

#define MOGP_KERNEL_RBF 0
#define MOGP_KERNEL_LINEAR 1

typedef struct mogp_engine mogp_engine;

/* Per-cluster model structure = the keyword arguments of GPobs.__init__ (obsmodel.py:7-8). */
typedef struct mogp_model_spec {
  int32_t kernel;              /* MOGP_KERNEL_RBF | MOGP_KERNEL_LINEAR   (obsmodel.py:25,35) */
  int32_t mean_func;           /* NegLinear mean m(x) = -A x on/off      (obsmodel.py:44-49, neg_linear.py:22) */
  int32_t signal_variance_fix; /* kernel variance excluded from the refit (obsmodel.py:29,38) */
  int32_t noise_variance_fix;  /* noise variance excluded from the refit  (obsmodel.py:51) */
} mogp_model_spec;

/* ---- engine lifetime -------------------------------------------------------------- */
int mogp_engine_create(int32_t device, mogp_engine** out);
void mogp_engine_destroy(mogp_engine* h);
const char* mogp_last_error(const mogp_engine* h); /* h may be NULL: error of the failed create */
int mogp_version(void);
/* Number of kernels this engine has launched since creation (bench.py's gpu_launches). */
int64_t mogp_launch_count(const mogp_engine* h);
/* Host waits of this engine sleep on an event (cudaEventBlockingSync) instead of spinning: for worker engines
   that share the host cores with many other threads.  Default off (lowest latency). */
int mogp_engine_set_blocking_sync(mogp_engine* h, int32_t on);
/* Block until the engine's stream is idle. */
int mogp_sync(mogp_engine* h);
/* Device-side time, in ms, spent between the two most recent mogp_timer_start/stop marks
   (CUDA events recorded on the engine's stream). */
int mogp_timer_start(mogp_engine* h);
int mogp_timer_stop(mogp_engine* h, double* ms_out);

/* ---- patients: the NaN-compacted rows of X, Y (mogp_constrained.py:246-247) --------- */
/* CSR: visits of patient n are x[off[n]..off[n+1]).  y_thr[n] is the RAW cell
   Y[n, index_to_eval] used by the monotonicity threshold (mogp_constrained.py:256-262);
   it may be NaN (then the test is false, as in NumPy).  y_thr may be NULL. */
int mogp_set_patients(mogp_engine* h, int64_t n_patients, const int64_t* off, const double* x,
                      const double* y, const double* y_thr);

/* Re-send the same table (same offsets) from host memory without disturbing an attached chain. */
int mogp_refresh_patients(mogp_engine* h, int64_t n_patients, const int64_t* off, const double* x, const double* y);
/* The same table built ON THE DEVICE from the reference's own data format: X, Y are the NaN-padded n_patients x T
   row-major matrices a caller hands to MoGP_constrained (mogp_constrained.py:28; rows compacted per patient at
   :246-247); thr_col = index_to_eval (1 with an onset anchor, :256).  A row whose X and Y have different numbers of
   non-NaN cells is MOGP_EINVAL.  Optional outputs (host, caller-allocated: off_out[n_patients + 1], x_out / y_out
   [n_patients * T]) receive the compacted table, *n_visits_out its length. */
int mogp_set_patients_matrix(mogp_engine* h, int64_t n_patients, int32_t T, const double* X, const double* Y, int32_t thr_col,
                             int64_t* off_out, double* x_out, double* y_out, int64_t* n_visits_out);

/* ---- accounting for bench.py ------------------------------------------------------------------ */
/* Bytes this engine has copied host->device / device->host since creation. */
int mogp_transfer_bytes(mogp_engine* h, int64_t* h2d, int64_t* d2h);
/* Device time of the dominant kernel (the scoring product T = M Kx): CUDA events on the engine's stream
   around every launch while enabled.  *ms = summed launch durations, *alg_bytes / *alg_flops = the
   algorithmic traffic / work of those launches (per (patient, cluster) pair
   8 (m^2/2 + 3m) + 16 n + 8 bytes and n m (m + 10) flops). */
int mogp_profile(mogp_engine* h, int32_t enable);
int mogp_profile_read(mogp_engine* h, int64_t* launches, double* ms, double* alg_bytes, double* alg_flops, int32_t reset);
/* Look-ahead accounting of the attached chain since mogp_chain_init: out[0] = window passes (one pass over every
   active factor), out[1] = re-scoring passes (clusters that changed), out[2] = patients served from windows,
   out[3] = (patient, cluster) pairs scored by the profiled launches since the last profile reset,
   out[4..6] = host seconds spent waiting for batches / re-scoring / committing moves (each includes its waits),
   out[7..10] = launches, device ms, algorithmic bytes and flops of the profiled RE-SCORING launches of the scoring
   product since the last profile reset (mogp_profile_read reports the whole passes only),
   out[11] = moves whose effect on the cached scores was a rank-n correction, out[12] = host seconds in them,
   out[13] = (cluster, batch) corrections replayed for the next batch's positions,
   out[14] = with MOGP_CHECK_CORRECTIONS=1 in the environment: worst relative disagreement between a rank-n corrected
   score / threshold mean and the same pair re-scored from the updated factor, out[15] = pairs checked,
   out[16] = (patient, cluster) pairs of the window passes, out[17] = those the monotonicity rule excluded on the device
   before the O(m^2) product (their mean alone decides, mogp_constrained.py:254-266),
   out[18] = host seconds waiting, when a batch starts, for the corrections chained behind its pass, out[19] = how many
   corrections were chained behind passes in flight, out[20] = device ms during which at least one profiled whole pass was
   running (the union of their intervals: two passes can be in flight; mogp_profile_read's ms is the per-launch sum) since
   the last profile reset - read it BEFORE mogp_profile_read(reset). */
int mogp_chain_counters(mogp_engine* h, double* out, int32_t n_out);

/* ---- one cluster's GP = one GPobs object ------------------------------------------------- */
/* GPobs.__init__ (obsmodel.py:7-61) minus data: allocates a cluster slot. */
int mogp_cluster_create(mogp_engine* h, const mogp_model_spec* spec, const double theta[4], int32_t* slot_out);
int mogp_cluster_free(mogp_engine* h, int32_t slot);
int mogp_cluster_set_theta(mogp_engine* h, int32_t slot, const double theta[4]); /* re-runs inference if data is set */
int mogp_cluster_get_theta(mogp_engine* h, int32_t slot, double theta[4]);
/* GPy set_XY (obsmodel.py:73,92) / GPRegression construction (:46): replace the data and run
   exact inference (fused K build + blocked Cholesky + inverse factor + alpha + log marginal). */
int mogp_cluster_set_data(mogp_engine* h, int32_t slot, int64_t m, const double* x, const double* y);
/* Same, data gathered on the device from the patient table (members in the given order). */
int mogp_cluster_set_members(mogp_engine* h, int32_t slot, int64_t n_members, const int64_t* patient_ids);
/* Append points at fixed theta (obsmodel.py:89-92 before optimize): rank-n update of the factors. */
int mogp_cluster_append(mogp_engine* h, int32_t slot, int64_t n, const double* x, const double* y);
/* Delete rows [p, p+n) of the cluster's data at fixed theta (the inverse of append; replaces the set_XY of
   update_data, obsmodel.py:68-73): closed-form re-triangularisation of the factor, chunks of <= 16 rows. */
int mogp_cluster_remove_rows(mogp_engine* h, int32_t slot, int64_t p, int64_t n);
/* Score of the cluster's own rows [p, p+n) (n <= 16) under the cluster WITHOUT them — what
   update_data + calc_score / get_pred_at_loc (mogp_constrained.py:239-266) compute — by the leave-block-out
   identity, leaving the cluster untouched.  *mu_thr = predictive mean at row p + thr_index (NaN if < 0). */
int mogp_cluster_leave_out_score(mogp_engine* h, int32_t slot, int64_t p, int64_t n, int32_t thr_index,
                                 double* score, double* mu_thr);
/* Diagnostics: the cluster's 32-double device stat block ([0..3] LML, r'Ky^-1 r, log det, x.alpha; [8] info;
   [10..13] dLML/dtheta; [14..15] cycle counts of the diagonal-tile Cholesky / inverse; [16..] theta). */
int mogp_cluster_debug_stat(mogp_engine* h, int32_t slot, double* out32);
int64_t mogp_cluster_size(mogp_engine* h, int32_t slot);
/* Diagnostics: the m x m inverse Cholesky factor M = chol(K + (noise + 1e-8) I)^-1 (row-major, lower) and
   alpha = Ky^-1 (y - m(x)) as currently held on the device.  Either output may be NULL. */
int mogp_cluster_get_factor(mogp_engine* h, int32_t slot, double* M_out, double* alpha_out);
int mogp_cluster_get_data(mogp_engine* h, int32_t slot, double* x, double* y);
/* model.log_likelihood() (obsmodel.py:56,65): log marginal likelihood, no prior term. */
int mogp_cluster_loglik(mogp_engine* h, int32_t slot, double* lml);
/* paramz Model._objective_grads (reached from obsmodel.py:94): f = -(LML + log prior + log Jacobian)
   and df/dt at optimiser coordinates t (Logexp space, free parameters only, GPy order).
   Leaves the cluster at theta(t), like paramz does.  *n_free_out = number of free parameters. */
int mogp_cluster_objective(mogp_engine* h, int32_t slot, const double* t, int32_t n_t, double* f, double* g);
/* The same evaluation for a cluster of `owner`, run on `worker`'s stream and workspaces (same device).  The
   refits of different clusters are independent (obsmodel.py:94 is per cluster): a caller may evaluate several
   clusters concurrently, one host thread + one worker engine each, provided the owner's stream is idle
   (mogp_sync) and no two threads touch the same slot. */
int mogp_cluster_objective_on(mogp_engine* worker, mogp_engine* owner, int32_t slot, const double* t, int32_t n_t,
                              double* f, double* g);
/* model.optimize() for several clusters at once (obsmodel.py:94 per cluster; the end-of-sweep refit of refit='sweep', or
   one cluster after a move): each listed cluster is re-optimised exactly as its own sequential optimize() would -
   L-BFGS-B (the engine's restatement of SciPy's, mogp_lbfgsb_*) on mogp_cluster_objective, warm-started at its current
   hyper-parameters, then f(x_opt) and the final set - but the evaluations of different clusters are in flight together:
   one lane (stream + factorisation workspace + captured graph) per cluster, at most max_concurrent of them (0 = default,
   32), all driven by the calling thread.  n_evals_out[i] (optional) = objective evaluations cluster i asked for (what
   paramz counts).  A cluster whose covariance stays non-positive-definite through GPy's jitter ladder 11 times in a row
   ends the call with MOGP_ENOTPD (paramz re-raises the LinAlgError after 10 swallowed failures). */
int mogp_refit_clusters(mogp_engine* h, int32_t n_slots, const int32_t* slots, int32_t max_concurrent, int32_t* n_evals_out);
/* Objective evaluations requested by mogp_refit_clusters since creation / those that ran on the device (the rest hit the
   "same hyper-parameters as the last evaluation" shortcut). */
int mogp_refit_counters(mogp_engine* h, int64_t* evaluations, int64_t* launched);
int mogp_cluster_n_free(mogp_engine* h, int32_t slot);
/* Current optimiser coordinates t = Logexp^-1(theta) of the free parameters. */
int mogp_cluster_get_optimizer_array(mogp_engine* h, int32_t slot, double* t, int32_t* n_t);
/* Gradient of the log marginal w.r.t. natural theta (4 entries, GPy order), for tests. */
int mogp_cluster_lml_grad(mogp_engine* h, int32_t slot, double grad_theta[4]);
/* GP._raw_predict + likelihood noise (obsmodel.py:85 -> model.predict without normalizer):
   mean[i], var[i] (= v* + noise) at xs[i]. */
int mogp_cluster_predict(mogp_engine* h, int32_t slot, int64_t n, const double* xs, double* mean, double* var);
/* model.log_predictive_density (obsmodel.py:80): per-visit log densities at (xs, ys). */
int mogp_cluster_log_predictive_density(mogp_engine* h, int32_t slot, int64_t n, const double* xs,
                                        const double* ys, double* lpd);

/* ---- batched pair scoring: GPobs.calc_score for patients x clusters (a1 + a2) ------------- */
/* Query patients given on the host as CSR.  score[p * n_slots + c] = sum over visits of the
   predictive log density of patient p under cluster slots[c] (obsmodel.py:75-81).
   mu_thr (optional) receives the predictive mean at visit `thr_index` of each patient
   (get_pred_at_loc, obsmodel.py:83-85), mu_thr[p * n_slots + c]. */
int mogp_score_pairs(mogp_engine* h, int64_t n_patients, const int64_t* off, const double* x, const double* y,
                     int32_t n_slots, const int32_t* slots, int32_t thr_index, double* score, double* mu_thr);
/* Same with every array already resident on the device (bench `value` path). */
int mogp_score_pairs_dev(mogp_engine* h, int64_t n_patients, int64_t n_visits, const int64_t* off_dev,
                         const double* x_dev, const double* y_dev, int32_t n_slots, const int32_t* slots,
                         int32_t thr_index, double* score_dev, double* mu_thr_dev);
/* New-cluster marginal (GPobs.__init__ on one patient, obsmodel.py:56; mogp_constrained.py:246-250,
   :330-339): ll[p] = log N(y_p | -A_p x_p, K_theta + (noise + 1e-8) I) with A_p = |a_draw[p]|. */
int mogp_new_cluster_ll(mogp_engine* h, const mogp_model_spec* spec, const double theta[4], int64_t n_patients,
                        const int64_t* off, const double* x, const double* y, const double* a_draw, double* ll);

/* ---- the sampler's initial partition (mogp_constrained.py:182-207) ------------------------------------------------ */
/* KMeans(n_clusters=K, random_state=seed, n_init=n_init).fit(x.reshape(-1, 1)).labels_ for the one-dimensional data the
   reference clusters (each patient's first-visit time, X[:, 1] with an onset anchor, :198-205), on the device: data
   centring, k-means++ seeding and Lloyd iterations with scikit-learn's float64 arithmetic and stopping rules (tol is the
   relative tolerance, 1e-4 in sklearn; max_iter 300), the best of n_init runs by inertia.  u[n_u] = the uniforms sklearn
   would take from RandomState(seed), i.e. RandomState(seed).random_sample(n_init * (1 + (K - 1) * (2 + int(ln K)))):
   like every other draw of the chain they are made on the host.  Outputs: labels_out[n] (cluster id per point, centres
   numbered in the order the seeding chose them - sklearn's numbering), optional centers_out[K], *inertia_out,
   *best_run_out, *n_iter_out. */
int mogp_kmeans_init(mogp_engine* h, int64_t n, const double* x, int32_t K, int32_t n_init, int32_t max_iter, double tol,
                     const double* u, int64_t n_u, int32_t* labels_out, double* centers_out, double* inertia_out,
                     int32_t* best_run_out, int32_t* n_iter_out);

/* ---- collapsed Gibbs sweep (mogp_constrained.py:231-284 + allocmodel.py:22-34) ------------- */
typedef struct mogp_chain_config {
  mogp_model_spec spec;
  double alpha;           /* DP concentration (allocmodel.py:8) */
  double signal_variance; /* initial theta of a new cluster (obsmodel.py:7-8, 27, 37) */
  double noise_variance;
  double lengthscale;     /* 4.0 in the reference */
  int32_t use_threshold;  /* threshold is not None (mogp_constrained.py:254) */
  int32_t thr_index;      /* 1 with onset anchor else 0 (mogp_constrained.py:256) */
  double threshold;
  /* 0: reference-exact data order — a leaving patient's cluster is re-extracted in patient-index order and
        refactorised (mogp_constrained.py:239-242), a joining patient's rows go last (obsmodel.py:77-78);
     1: rank-n up/down-dates at fixed theta + leave-block-out score for the patient's own cluster
        (same mathematics, rows kept in arrival order; agrees with mode 0 to rounding). */
  int32_t fast_updates;
  /* mogp_chain_sweep, fast mode: the sweep order is known in advance, so consecutive patients totalling about
     `lookahead` visits (query columns; the cut is placed where the last group of 8 columns is fullest) are scored against every active cluster in ONE pass over the factors; the
     chain stays sequential (a cached score is used only while its cluster is unchanged, the rest is re-scored).
     0 = engine default (64), negative = off (one pass per patient), capped at 64.  While the host draws and moves the
     patients of one batch, the pass of the next batch already runs on a low-priority stream (MOGP_PIPELINE=1 in the
     environment turns that overlap off; MOGP_LOOKAHEAD overrides this field: tuning and tests). */
  int32_t lookahead;
} mogp_chain_config;

/* initialize_sampler + calc_suffstats (mogp_constrained.py:187-214,220): z[N] = initial cluster id per patient
   (ids 0..n_ids-1), theta_of_id[4*k..] = hyper-parameters of cluster k (A = |randn| drawn by the caller in id
   order).  Builds one cluster slot per non-empty id from the patient table (members in index order). */
int mogp_chain_init(mogp_engine* h, const mogp_chain_config* cfg, int64_t n_patients, const int32_t* z,
                    int32_t n_ids, const double* theta_of_id);
/* One patient step in three phases, so that a caller can teacher-force or refit between them:
   begin  : take patient n out of its cluster (mogp_constrained.py:232-242, allocmodel.py:24);
   score  : score[i] = log prior weight + log likelihood of active id ids[i] (last = new cluster), with the
            monotonicity rule applied (:244-271); a_draw = the randn consumed by the newcomer's NegLinear map;
   commit : assign to chosen_id (allocmodel.py:29-34; :279-284 WITHOUT the hyper-parameter refit — the caller
            runs it through mogp_cluster_objective when the schedule asks for one). */
int mogp_chain_step_begin(mogp_engine* h, int64_t n);
int mogp_chain_step_score(mogp_engine* h, double a_draw, int32_t* n_active, int32_t* ids_out, double* score_out,
                          int32_t cap_out);
int mogp_chain_step_commit(mogp_engine* h, int32_t chosen_id, double a_draw, int32_t* is_new);
/* begin + score + draw + commit with pre-drawn randomness (u = the uniform consumed by np.random.choice, :277).
   Outputs (optional): chosen id, p_out[0..*n_active) = probabilities over ids_out. */
int mogp_chain_step(mogp_engine* h, int64_t n, double a_draw, double u, int32_t* chosen_id, int32_t* is_new,
                    int32_t* n_active, int32_t* ids_out, double* p_out, int32_t cap_out);
/* n_steps consecutive steps (one sweep when n_steps = N) without refits: order[], a_draw[], u[] per step.
   min_margin (optional) = the smallest |u - cdf edge| seen (assignment-parity diagnostic). */
int mogp_chain_sweep(mogp_engine* h, int64_t n_steps, const int64_t* order, const double* a_draw, const double* u,
                     int32_t* chosen_ids, double* min_margin);
int mogp_chain_get_state(mogp_engine* h, int32_t* z /*N*/, int64_t* Nk, int32_t* slot_of_id, int32_t cap_ids,
                         int32_t* n_ids);
/* Last step's minimum |u - cdf edge|. */
double mogp_chain_last_margin(const mogp_engine* h);
/* self.p[n] (mogp_constrained.py:276): with recording on, every step of mogp_chain_sweep / mogp_chain_step keeps its
   membership probabilities; mogp_chain_get_probs returns those of the steps since the last mogp_chain_sweep began: step
   s covers [off[s], off[s+1]) of ids[] (active ids in the reference's order, last = the new cluster) and p[].  With
   off = ids = p = NULL only *n_steps and *total are written (call twice). */
int mogp_chain_record_probs(mogp_engine* h, int32_t on);
int mogp_chain_get_probs(mogp_engine* h, int64_t* n_steps, int64_t* total, int64_t* off, int32_t* ids, double* p,
                         int64_t cap_steps, int64_t cap_total);
/* The draw alone (host arithmetic, no device): p = exp(score - logsumexp(score)) and NumPy's
   choice(): cdf = cumsum(p); cdf /= cdf[-1]; index = searchsorted(cdf, u, 'right')  (mogp_constrained.py:272-277). */
int mogp_host_choice(const double* score, int32_t n, double u, double* p_out, int32_t* index_out, double* margin_out);

/* ---- the optimiser of model.optimize() as a host-side state machine (no device) -------------------------------- */
/* paramz's default optimiser (obsmodel.py:94, mogp_constrained.py:333) is scipy.optimize.fmin_l_bfgs_b(f_fp, x0,
   maxfun=1000, maxiter=1000): L-BFGS-B 3.0 without bounds, m = 10, factr = 1e7, pgtol = 1e-5, maxls = 20.  The engine
   carries its own restatement (csrc/lbfgsb.hpp) so that the refits of many clusters can be driven by ONE host thread
   while their evaluations run on the device (mogp_refit_clusters); these entry points expose it in reverse-communication
   form for tests against SciPy.  mogp_lbfgsb_step: on entry *f, g[n] hold the objective at x when the previous call
   returned 1; returns 1 = evaluate (f, g) at x[n] and call again, 2 = new iterate (call again), 3 / 4 = converged
   (projected gradient <= pgtol / relative reduction <= factr * eps), 5 = abnormal termination in the line search,
   6 = iteration or evaluation limit. */
typedef struct mogp_lbfgsb mogp_lbfgsb;
int mogp_lbfgsb_create(int32_t n, int32_t m, double factr, double pgtol, int32_t maxls, const double* x0, mogp_lbfgsb** out);
int mogp_lbfgsb_step(mogp_lbfgsb* s, double* x, double* f, double* g);
void mogp_lbfgsb_destroy(mogp_lbfgsb* s);

#ifdef __cplusplus
}
#endif
#endif /* MOGP_B200_H */
