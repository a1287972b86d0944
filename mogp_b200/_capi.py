"""ctypes binding of include/mogp_b200.h — the only way the package reaches the device.

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible,
the calls raise.  (`oracle/` is test infrastructure and is never imported from here.)
"""
import ctypes as C
import os

import numpy as np

from ._build import LIB

MOGP_OK, MOGP_EINVAL, MOGP_ECUDA, MOGP_ENOMEM, MOGP_ENOTPD, MOGP_EPROB, MOGP_ESTATE = 0, -1, -2, -3, -4, -5, -6
KERNEL_RBF, KERNEL_LINEAR = 0, 1
KERNEL_CODE = {"rbf": KERNEL_RBF, "linear": KERNEL_LINEAR}


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mogp_b200 engine error {}: {}".format(code, msg))
        self.code = code


class ModelSpec(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("mean_func", C.c_int32), ("signal_variance_fix", C.c_int32),
                ("noise_variance_fix", C.c_int32)]


class ChainConfig(C.Structure):
    _fields_ = [("spec", ModelSpec), ("alpha", C.c_double), ("signal_variance", C.c_double),
                ("noise_variance", C.c_double), ("lengthscale", C.c_double), ("use_threshold", C.c_int32),
                ("thr_index", C.c_int32), ("threshold", C.c_double), ("fast_updates", C.c_int32),
                ("lookahead", C.c_int32)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_H = C.c_void_p

# name -> (restype, argtypes): every symbol include/mogp_b200.h declares
SIGNATURES = {
    "mogp_engine_create": (C.c_int, [C.c_int32, C.POINTER(_H)]),
    "mogp_engine_destroy": (None, [_H]),
    "mogp_last_error": (C.c_char_p, [_H]),
    "mogp_version": (C.c_int, []),
    "mogp_launch_count": (C.c_int64, [_H]),
    "mogp_engine_set_blocking_sync": (C.c_int, [_H, C.c_int32]),
    "mogp_sync": (C.c_int, [_H]),
    "mogp_timer_start": (C.c_int, [_H]),
    "mogp_timer_stop": (C.c_int, [_H, _dp]),
    "mogp_set_patients": (C.c_int, [_H, C.c_int64, _lp, _dp, _dp, _dp]),
    "mogp_refresh_patients": (C.c_int, [_H, C.c_int64, _lp, _dp, _dp]),
    "mogp_set_patients_matrix": (C.c_int, [_H, C.c_int64, C.c_int32, _dp, _dp, C.c_int32, _lp, _dp, _dp, _lp]),
    "mogp_transfer_bytes": (C.c_int, [_H, _lp, _lp]),
    "mogp_profile": (C.c_int, [_H, C.c_int32]),
    "mogp_profile_read": (C.c_int, [_H, _lp, _dp, _dp, _dp, C.c_int32]),
    "mogp_chain_counters": (C.c_int, [_H, _dp, C.c_int32]),
    "mogp_cluster_create": (C.c_int, [_H, C.POINTER(ModelSpec), _dp, _ip]),
    "mogp_cluster_free": (C.c_int, [_H, C.c_int32]),
    "mogp_cluster_set_theta": (C.c_int, [_H, C.c_int32, _dp]),
    "mogp_cluster_get_theta": (C.c_int, [_H, C.c_int32, _dp]),
    "mogp_cluster_set_data": (C.c_int, [_H, C.c_int32, C.c_int64, _dp, _dp]),
    "mogp_cluster_set_members": (C.c_int, [_H, C.c_int32, C.c_int64, _lp]),
    "mogp_cluster_append": (C.c_int, [_H, C.c_int32, C.c_int64, _dp, _dp]),
    "mogp_cluster_remove_rows": (C.c_int, [_H, C.c_int32, C.c_int64, C.c_int64]),
    "mogp_cluster_debug_stat": (C.c_int, [_H, C.c_int32, _dp]),
    "mogp_cluster_size": (C.c_int64, [_H, C.c_int32]),
    "mogp_cluster_get_factor": (C.c_int, [_H, C.c_int32, _dp, _dp]),
    "mogp_cluster_get_data": (C.c_int, [_H, C.c_int32, _dp, _dp]),
    "mogp_cluster_loglik": (C.c_int, [_H, C.c_int32, _dp]),
    "mogp_cluster_objective": (C.c_int, [_H, C.c_int32, _dp, C.c_int32, _dp, _dp]),
    "mogp_cluster_objective_on": (C.c_int, [_H, _H, C.c_int32, _dp, C.c_int32, _dp, _dp]),
    "mogp_refit_clusters": (C.c_int, [_H, C.c_int32, _ip, C.c_int32, _ip]),
    "mogp_refit_counters": (C.c_int, [_H, _lp, _lp]),
    "mogp_cluster_n_free": (C.c_int, [_H, C.c_int32]),
    "mogp_cluster_get_optimizer_array": (C.c_int, [_H, C.c_int32, _dp, _ip]),
    "mogp_cluster_lml_grad": (C.c_int, [_H, C.c_int32, _dp]),
    "mogp_cluster_predict": (C.c_int, [_H, C.c_int32, C.c_int64, _dp, _dp, _dp]),
    "mogp_cluster_log_predictive_density": (C.c_int, [_H, C.c_int32, C.c_int64, _dp, _dp, _dp]),
    "mogp_cluster_leave_out_score": (C.c_int, [_H, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _dp, _dp]),
    "mogp_score_pairs": (C.c_int, [_H, C.c_int64, _lp, _dp, _dp, C.c_int32, _ip, C.c_int32, _dp, _dp]),
    "mogp_score_pairs_dev": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, _ip,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "mogp_new_cluster_ll": (C.c_int, [_H, C.POINTER(ModelSpec), _dp, C.c_int64, _lp, _dp, _dp, _dp, _dp]),
    "mogp_kmeans_init": (C.c_int, [_H, C.c_int64, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_double, _dp, C.c_int64, _ip, _dp,
                                   _dp, _ip, _ip]),
    "mogp_chain_init": (C.c_int, [_H, C.POINTER(ChainConfig), C.c_int64, _ip, C.c_int32, _dp]),
    "mogp_chain_step_begin": (C.c_int, [_H, C.c_int64]),
    "mogp_chain_step_score": (C.c_int, [_H, C.c_double, _ip, _ip, _dp, C.c_int32]),
    "mogp_chain_step_commit": (C.c_int, [_H, C.c_int32, C.c_double, _ip]),
    "mogp_chain_step": (C.c_int, [_H, C.c_int64, C.c_double, C.c_double, _ip, _ip, _ip, _ip, _dp, C.c_int32]),
    "mogp_chain_sweep": (C.c_int, [_H, C.c_int64, _lp, _dp, _dp, _ip, _dp]),
    "mogp_chain_get_state": (C.c_int, [_H, _ip, _lp, _ip, C.c_int32, _ip]),
    "mogp_chain_last_margin": (C.c_double, [_H]),
    "mogp_chain_record_probs": (C.c_int, [_H, C.c_int32]),
    "mogp_chain_get_probs": (C.c_int, [_H, _lp, _lp, _lp, _ip, _dp, C.c_int64, C.c_int64]),
    "mogp_host_choice": (C.c_int, [_dp, C.c_int32, C.c_double, _dp, _ip, _dp]),
    "mogp_lbfgsb_create": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, _dp, C.POINTER(_H)]),
    "mogp_lbfgsb_step": (C.c_int, [_H, _dp, _dp, _dp]),
    "mogp_lbfgsb_destroy": (None, [_H]),
}

_lib = None


def load_library():
    """dlopen the in-tree library (built by `__graft_entry__.build()` / `python -m mogp_b200._build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise ImportError("mogp_b200: {} is missing — build it with `python -m mogp_b200._build` "
                          "(there is no CPU fallback)".format(LIB))
    lib = C.CDLL(LIB)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def lptr(a):
    return a.ctypes.data_as(_lp) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """One engine handle = one CUDA stream + cluster slots (see include/mogp_b200.h)."""

    def __init__(self, device=None, blocking_sync=False):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = _H()
        rc = self.lib.mogp_engine_create(int(device), C.byref(h))
        if rc != MOGP_OK:
            raise EngineError(rc, self.lib.mogp_last_error(None).decode())
        self.h = h
        self.device = int(device)
        if blocking_sync:
            self.lib.mogp_engine_set_blocking_sync(self.h, 1)

    def close(self):
        if getattr(self, "h", None):
            self.lib.mogp_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != MOGP_OK:
            code = rc
            msg = self.lib.mogp_last_error(self.h).decode()
            if code == MOGP_ENOTPD:
                raise np.linalg.LinAlgError(msg)        # what GPy's jitchol raises
            if code == MOGP_EPROB:
                raise ValueError("probabilities contain NaN")   # what np.random.choice raises
            raise EngineError(code, msg)

    # -- convenience wrappers ---------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.mogp_launch_count(self.h))

    def sync(self):
        self.check(self.lib.mogp_sync(self.h))

    def set_patients(self, off, x, y, y_thr=None):
        off = np.ascontiguousarray(off, dtype=np.int64)
        x, y = f64(x), f64(y)
        y_thr = None if y_thr is None else f64(y_thr)
        self.check(self.lib.mogp_set_patients(self.h, len(off) - 1, lptr(off), dptr(x), dptr(y), dptr(y_thr)))

    def set_patients_matrix(self, X, Y, thr_col=-1):
        """Patient table from the reference's NaN-padded N x T matrices, compacted on the device; returns the CSR
        (off, x, y) it built (== utils.pack_patients(X, Y), bit for bit)."""
        X, Y = f64(X), f64(Y)
        if X.ndim != 2 or X.shape != Y.shape:
            raise ValueError("X and Y must be N x T matrices of the same shape")
        N, T = X.shape
        off = np.zeros(N + 1, dtype=np.int64)
        x, y = np.empty(N * T), np.empty(N * T)
        nv = C.c_int64()
        rc = self.lib.mogp_set_patients_matrix(self.h, N, T, dptr(X), dptr(Y), int(thr_col), lptr(off), dptr(x), dptr(y),
                                               C.cast(C.byref(nv), _lp))
        if rc == MOGP_EINVAL and b"different numbers" in self.lib.mogp_last_error(self.h):
            raise ValueError("X and Y have different numbers of non-NaN cells in some row")
        self.check(rc)
        return off, x[:nv.value].copy(), y[:nv.value].copy()

    def refresh_patients(self, off, x, y):
        self.check(self.lib.mogp_refresh_patients(self.h, len(off) - 1, lptr(off), dptr(x), dptr(y)))

    def transfer_bytes(self):
        a, b = C.c_int64(), C.c_int64()
        self.check(self.lib.mogp_transfer_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile(self, on):
        self.check(self.lib.mogp_profile(self.h, int(bool(on))))

    def profile_read(self, reset=True):
        n, ms, by, fl = C.c_int64(), C.c_double(), C.c_double(), C.c_double()
        self.check(self.lib.mogp_profile_read(self.h, C.byref(n), C.byref(ms), C.byref(by), C.byref(fl), int(reset)))
        return dict(launches=n.value, ms=ms.value, alg_bytes=by.value, alg_flops=fl.value)

    def chain_counters(self):
        out = np.zeros(21)
        self.check(self.lib.mogp_chain_counters(self.h, dptr(out), 21))
        return dict(window_passes=int(out[0]), rescoring_passes=int(out[1]), window_patients=int(out[2]), pairs=out[3],
                    open_s=out[4], rescore_s=out[5], commit_s=out[6], rescoring_launches=int(out[7]), rescoring_ms=out[8],
                    rescoring_bytes=out[9], rescoring_flops=out[10], rank_corrections=int(out[11]), correct_s=out[12], replayed=int(out[13]),
                    max_correction_err=out[14], corrections_checked=int(out[15]), window_pairs=int(out[16]),
                    pruned_pairs=int(out[17]), chained_wait_s=out[18], chained_corrections=int(out[19]),
                    passes_union_ms=out[20])

    def timer_start(self):
        self.check(self.lib.mogp_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self.check(self.lib.mogp_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def cluster_create(self, spec, theta):
        slot = C.c_int32(-1)
        theta = f64(theta)
        self.check(self.lib.mogp_cluster_create(self.h, C.byref(spec), dptr(theta), C.byref(slot)))
        return slot.value

    def cluster_free(self, slot):
        self.check(self.lib.mogp_cluster_free(self.h, slot))

    def cluster_set_data(self, slot, x, y):
        x, y = f64(x).ravel(), f64(y).ravel()
        self.check(self.lib.mogp_cluster_set_data(self.h, slot, x.size, dptr(x), dptr(y)))

    def cluster_set_members(self, slot, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self.check(self.lib.mogp_cluster_set_members(self.h, slot, ids.size, lptr(ids)))

    def cluster_append(self, slot, x, y):
        x, y = f64(x).ravel(), f64(y).ravel()
        self.check(self.lib.mogp_cluster_append(self.h, slot, x.size, dptr(x), dptr(y)))

    def cluster_remove_rows(self, slot, p, n):
        self.check(self.lib.mogp_cluster_remove_rows(self.h, slot, int(p), int(n)))

    def cluster_size(self, slot):
        return int(self.lib.mogp_cluster_size(self.h, slot))

    def cluster_get_data(self, slot):
        m = self.cluster_size(slot)
        x, y = np.empty(m), np.empty(m)
        self.check(self.lib.mogp_cluster_get_data(self.h, slot, dptr(x), dptr(y)))
        return x, y

    def cluster_debug_stat(self, slot):
        out = np.empty(32)
        self.check(self.lib.mogp_cluster_debug_stat(self.h, slot, dptr(out)))
        return out

    def cluster_get_factor(self, slot):
        m = self.cluster_size(slot)
        M, alpha = np.empty((m, m)), np.empty(m)
        self.check(self.lib.mogp_cluster_get_factor(self.h, slot, dptr(M), dptr(alpha)))
        return np.tril(M), alpha      # tiles above the diagonal are never written on the device

    def cluster_get_theta(self, slot):
        th = np.empty(4)
        self.check(self.lib.mogp_cluster_get_theta(self.h, slot, dptr(th)))
        return th

    def cluster_set_theta(self, slot, theta):
        theta = f64(theta)
        self.check(self.lib.mogp_cluster_set_theta(self.h, slot, dptr(theta)))

    def cluster_loglik(self, slot):
        v = C.c_double()
        self.check(self.lib.mogp_cluster_loglik(self.h, slot, C.byref(v)))
        return v.value

    def cluster_lml_grad(self, slot):
        g = np.empty(4)
        self.check(self.lib.mogp_cluster_lml_grad(self.h, slot, dptr(g)))
        return g

    def cluster_n_free(self, slot):
        return int(self.lib.mogp_cluster_n_free(self.h, slot))

    def cluster_optimizer_array(self, slot):
        t = np.empty(4)
        n = C.c_int32()
        self.check(self.lib.mogp_cluster_get_optimizer_array(self.h, slot, dptr(t), C.byref(n)))
        return t[:n.value].copy()

    def refit_clusters(self, slots, max_concurrent=0):
        """model.optimize() of every listed cluster (obsmodel.py:94), their evaluations in flight together, driven by this
        thread (mogp_refit_clusters).  Returns the objective evaluations each cluster asked for."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        n_evals = np.zeros(slots.size, dtype=np.int32)
        if slots.size:
            self.check(self.lib.mogp_refit_clusters(self.h, slots.size, iptr(slots), int(max_concurrent), iptr(n_evals)))
        return n_evals

    def refit_counters(self):
        a, b = C.c_int64(), C.c_int64()
        self.check(self.lib.mogp_refit_counters(self.h, C.byref(a), C.byref(b)))
        return dict(evaluations=a.value, launched=b.value)

    def cluster_objective(self, slot, t=None, want_grad=True, worker=None, nfree=None):
        """(f, df/dt) of cluster `slot`; with `worker` (another Engine on the same device) the evaluation runs
        on the worker's stream and workspaces, so several clusters can be refitted concurrently (`nfree` given: the
        owner engine is not touched from the worker's thread)."""
        f = C.c_double()
        if nfree is None:
            nfree = self.cluster_n_free(slot)
        g = np.empty(max(nfree, 1))
        if t is not None:
            t = f64(t)
        if worker is None:
            rc = self.lib.mogp_cluster_objective(self.h, slot, dptr(t), 0 if t is None else t.size, C.byref(f),
                                                 dptr(g) if want_grad else None)
            self.check(rc)
        else:
            rc = self.lib.mogp_cluster_objective_on(worker.h, self.h, slot, dptr(t), 0 if t is None else t.size,
                                                    C.byref(f), dptr(g) if want_grad else None)
            worker.check(rc)
        return f.value, g[:nfree].copy()

    def cluster_predict(self, slot, xs):
        xs = f64(xs).ravel()
        mean, var = np.empty(xs.size), np.empty(xs.size)
        self.check(self.lib.mogp_cluster_predict(self.h, slot, xs.size, dptr(xs), dptr(mean), dptr(var)))
        return mean, var

    def cluster_lpd(self, slot, xs, ys):
        xs, ys = f64(xs).ravel(), f64(ys).ravel()
        out = np.empty(xs.size)
        self.check(self.lib.mogp_cluster_log_predictive_density(self.h, slot, xs.size, dptr(xs), dptr(ys), dptr(out)))
        return out

    def cluster_leave_out_score(self, slot, p, n, thr_index=-1):
        s, mu = C.c_double(), C.c_double()
        self.check(self.lib.mogp_cluster_leave_out_score(self.h, slot, int(p), int(n), int(thr_index),
                                                         C.byref(s), C.byref(mu)))
        return s.value, mu.value

    def score_pairs(self, off, x, y, slots, thr_index=-1, want_mu=False):
        off = np.ascontiguousarray(off, dtype=np.int64)
        x, y = f64(x), f64(y)
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        P = len(off) - 1
        score = np.empty((P, slots.size))
        mu = np.empty((P, slots.size)) if want_mu else None
        self.check(self.lib.mogp_score_pairs(self.h, P, lptr(off), dptr(x), dptr(y), slots.size, iptr(slots),
                                             int(thr_index), dptr(score), dptr(mu)))
        return (score, mu) if want_mu else score

    def new_cluster_ll(self, spec, theta, off, x, y, a_draw):
        off = np.ascontiguousarray(off, dtype=np.int64)
        x, y, theta = f64(x), f64(y), f64(theta)
        a_draw = None if a_draw is None else f64(a_draw)
        out = np.empty(len(off) - 1)
        self.check(self.lib.mogp_new_cluster_ll(self.h, C.byref(spec), dptr(theta), len(off) - 1, lptr(off), dptr(x),
                                                dptr(y), dptr(a_draw), dptr(out)))
        return out

    def kmeans_init(self, x, n_clusters, seed, n_init=10, max_iter=300, tol=1e-4, details=False):
        """sklearn.cluster.KMeans(n_clusters, random_state=seed, n_init=n_init).fit(x[:, None]).labels_ on the device
        (mogp_constrained.py:182-207; 1-D data).  The uniforms sklearn's k-means++ would draw from RandomState(seed) are
        made here, in its order."""
        x = f64(x).ravel()
        K = int(n_clusters)
        trials = 2 + int(np.log(K))
        u = np.random.RandomState(seed).random_sample(int(n_init) * (1 + (K - 1) * trials))
        labels = np.empty(x.size, dtype=np.int32)
        centers = np.empty(K)
        inertia = C.c_double()
        best, it = C.c_int32(), C.c_int32()
        self.check(self.lib.mogp_kmeans_init(self.h, x.size, dptr(x), K, int(n_init), int(max_iter), float(tol), dptr(u),
                                             u.size, iptr(labels), dptr(centers), C.byref(inertia), C.byref(best),
                                             C.byref(it)))
        if details:
            return labels, dict(centers=centers, inertia=inertia.value, best_run=best.value, n_iter=it.value)
        return labels

    # -- chain ----------------------------------------------------------------------------------
    def chain_init(self, cfg, z, theta_of_id):
        z = np.ascontiguousarray(z, dtype=np.int32)
        theta_of_id = f64(theta_of_id).reshape(-1, 4)
        self.check(self.lib.mogp_chain_init(self.h, C.byref(cfg), z.size, iptr(z), theta_of_id.shape[0],
                                            dptr(theta_of_id)))

    def chain_step_begin(self, n):
        self.check(self.lib.mogp_chain_step_begin(self.h, int(n)))

    def chain_step_score(self, a_draw, cap=4096):
        ids = np.empty(cap, dtype=np.int32)
        sc = np.empty(cap)
        n = C.c_int32()
        self.check(self.lib.mogp_chain_step_score(self.h, float(a_draw), C.byref(n), iptr(ids), dptr(sc), cap))
        return ids[:n.value].copy(), sc[:n.value].copy()

    def chain_step_commit(self, chosen, a_draw):
        is_new = C.c_int32()
        self.check(self.lib.mogp_chain_step_commit(self.h, int(chosen), float(a_draw), C.byref(is_new)))
        return bool(is_new.value)

    def chain_step(self, n, a_draw, u, cap=4096):
        ids = np.empty(cap, dtype=np.int32)
        p = np.empty(cap)
        na, chosen, is_new = C.c_int32(), C.c_int32(), C.c_int32()
        self.check(self.lib.mogp_chain_step(self.h, int(n), float(a_draw), float(u), C.byref(chosen), C.byref(is_new),
                                            C.byref(na), iptr(ids), dptr(p), cap))
        return chosen.value, bool(is_new.value), ids[:na.value].copy(), p[:na.value].copy()

    def chain_sweep(self, order, a_draw, u):
        order = np.ascontiguousarray(order, dtype=np.int64)
        a_draw = None if a_draw is None else f64(a_draw)
        u = f64(u)
        chosen = np.empty(order.size, dtype=np.int32)
        mm = C.c_double()
        self.check(self.lib.mogp_chain_sweep(self.h, order.size, lptr(order), dptr(a_draw), dptr(u), iptr(chosen),
                                             C.byref(mm)))
        return chosen, mm.value

    def chain_record_probs(self, on=True):
        self.check(self.lib.mogp_chain_record_probs(self.h, int(bool(on))))

    def chain_probs(self):
        """Membership probabilities of every step of the last chain_sweep (needs chain_record_probs): off, ids, p."""
        ns, tot = C.c_int64(), C.c_int64()
        self.check(self.lib.mogp_chain_get_probs(self.h, C.byref(ns), C.byref(tot), None, None, None, 0, 0))
        off = np.zeros(ns.value + 1, dtype=np.int64)
        ids = np.zeros(max(tot.value, 1), dtype=np.int32)
        p = np.zeros(max(tot.value, 1))
        self.check(self.lib.mogp_chain_get_probs(self.h, C.byref(ns), C.byref(tot), lptr(off), iptr(ids), dptr(p),
                                                 ns.value, max(tot.value, 1)))
        return off, ids[:tot.value], p[:tot.value]

    def chain_state(self, N):
        n_ids = C.c_int32()
        self.check(self.lib.mogp_chain_get_state(self.h, None, None, None, 0, C.byref(n_ids)))
        k = n_ids.value
        z = np.empty(N, dtype=np.int32)
        Nk = np.empty(k, dtype=np.int64)
        slots = np.empty(k, dtype=np.int32)
        self.check(self.lib.mogp_chain_get_state(self.h, iptr(z), lptr(Nk), iptr(slots), k, C.byref(n_ids)))
        return z, Nk, slots


def host_choice(score, u):
    """The engine's host-side draw (no device needed)."""
    lib = load_library()
    score = f64(score)
    p = np.empty(score.size)
    idx, margin = C.c_int32(), C.c_double()
    rc = lib.mogp_host_choice(dptr(score), score.size, float(u), dptr(p), C.byref(idx), C.byref(margin))
    if rc == MOGP_EPROB:
        raise ValueError("probabilities contain NaN")
    if rc != MOGP_OK:
        raise EngineError(rc, "mogp_host_choice")
    return idx.value, p, margin.value


def lbfgsb_minimize(fun, x0, m=10, factr=1e7, pgtol=1e-5, maxls=20):
    """Drive the engine's L-BFGS-B state machine (csrc/lbfgsb.hpp) with a Python objective `fun(x) -> (f, g)`:
    the counterpart of scipy.optimize.fmin_l_bfgs_b(fun, x0) for tests.  Returns (x, f, info)."""
    lib = load_library()
    x = f64(x0).copy()
    n = x.size
    h = _H()
    rc = lib.mogp_lbfgsb_create(n, m, float(factr), float(pgtol), int(maxls), dptr(x), C.byref(h))
    if rc != MOGP_OK:
        raise EngineError(rc, "mogp_lbfgsb_create")
    f = C.c_double(0.0)
    g = np.zeros(n)
    evals, iters, xs = 0, 0, []
    try:
        while True:
            task = lib.mogp_lbfgsb_step(h, dptr(x), C.byref(f), dptr(g))
            if task == 1:
                fv, gv = fun(x.copy())
                xs.append(x.copy())
                f.value = float(fv)
                g[:] = np.asarray(gv, dtype=np.float64)
                evals += 1
            elif task == 2:
                iters += 1
            else:
                break
    finally:
        lib.mogp_lbfgsb_destroy(h)
    return x.copy(), f.value, dict(task=task, funcalls=evals, nit=iters, points=xs, grad=g.copy())


_default_engine = None


def default_engine():
    """Process-wide engine used by GPobs objects created without an explicit one."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine
