"""Post-processing and forecasting on the engine — the callers just downstream of the likelihood path
(SURVEY.md §8f rows 2-3): `analysis/model_postprocessing.py:6-104` (monotonicity check, splitting of non-monotone
clusters) and `analysis/analysis_utils.py:80-92,493-521` (cluster-mean forecast of a patient, RMSE against the
ranked cluster).  Same function names, arguments and results; the per-window / per-patient Python loops of the
reference become one batched device call each.
"""
import numpy as np

from .obsmodel import GPobs


def _window_starts(lo, hi, window, step=0.1):
    """The reference's `min_point += 0.1` walk (model_postprocessing.py:23-29), same floating-point accumulation."""
    pts = []
    p = lo
    while (p + window) < hi:
        pts.append(p)
        p += step
    return np.array(pts, dtype=np.float64)


def find_cluster_jumps(obs, thresh=10, window=1):
    """True if no sliding window [t, t+window] of the cluster's mean rises by more than `thresh`
    (model_postprocessing.py:21-32). One predict call over all window edges instead of one per window."""
    X = np.asarray(obs.X, dtype=np.float64)
    starts = _window_starts(np.min(X), np.max(X), window)
    if starts.size == 0:
        return True
    mean, _ = obs.model.predict(np.concatenate([starts, starts + window]).reshape(-1, 1))
    mean = mean.ravel()
    rise = mean[starts.size:] - mean[:starts.size]
    return not bool(np.any(rise > thresh))


def check_model_monotonicity(model, thresh=10, window=1, num_obs=20):
    """model_postprocessing.py:6-19: all of the `num_obs` largest clusters are jump-free."""
    nc = len(np.where(model.allocmodel.Nk > 0)[0])
    idx = np.argsort(-model.allocmodel.Nk)[0:nc]
    if num_obs is not None:
        idx = idx[:num_obs]
    flag = True
    for k in idx:
        flag = find_cluster_jumps(obs=model.obsmodel[k], thresh=thresh, window=window) and flag
    return flag


def generate_clust_obs(model, clust, optimize_params=True):
    """model_postprocessing.py:76-104: fit a fresh GP to the patients `clust` and append it as a new cluster."""
    xclust, yclust = model.X[clust], model.Y[clust]
    new_obs = GPobs(xclust[~np.isnan(xclust)].reshape(-1, 1), yclust[~np.isnan(yclust)].reshape(-1, 1),
                    kernel=model.kernel, mean_func=model.mean_func, signal_variance=model.signal_variance,
                    signal_variance_fix=model.signal_variance_fix, noise_variance=model.noise_variance,
                    noise_variance_fix=model.noise_variance_fix, engine=getattr(model, 'engine', None))
    if optimize_params:
        new_obs.model.optimize()
    model.allocmodel.Nk = np.hstack([model.allocmodel.Nk, xclust.shape[0]])
    new_comp = np.where(model.allocmodel.Nk > 0)[0][-1]
    model.z[clust] = new_comp
    model.obsmodel[new_comp] = new_obs


def split_nonmonotonic_clusters(model, rand_seed=0, thresh=10):
    """model_postprocessing.py:48-73: clusters whose mean jumps are split in two by k-means on the first real visit
    time, each half refitted; the old cluster id is retired."""
    from sklearn.cluster import KMeans
    nc = len(np.where(model.allocmodel.Nk > 0)[0])
    idx = np.argsort(-model.allocmodel.Nk)[0:nc]
    for k in idx:
        if find_cluster_jumps(obs=model.obsmodel[k], thresh=thresh) is False:
            ind_z = np.where(model.z == k)[0]
            labels = np.array(KMeans(n_clusters=2, random_state=rand_seed).fit(model.X[ind_z][:, 1].reshape(-1, 1)).labels_)
            model.obsmodel[k] = None
            model.allocmodel.Nk[k] = 0
            generate_clust_obs(model, ind_z[np.where(labels == 0)])
            generate_clust_obs(model, ind_z[np.where(labels == 1)])
    model._apply_normalizer()


# ---- forecasting (analysis/analysis_utils.py:80-92, 493-521) ---------------------------------------------------
def calc_y_pred_model(mod, x_real, cur_clust):
    """Mean trajectory of cluster `cur_clust` at the patient's visit times (analysis_utils.py:80-84)."""
    return mod.obsmodel[cur_clust].model.predict(np.asarray(x_real, dtype=np.float64).reshape(-1, 1))[0].transpose()


def rmse(predictions, targets):
    return np.sqrt(((predictions - targets) ** 2).mean())


def calc_error(y_real, y_pred):
    return rmse(y_pred, y_real) if len(y_real) > 0 else np.nan


def rank_and_forecast_errors(model, XA, YA, a_draws=None):
    """The body of `gen_df_err_extend` (analysis_utils.py:493-521) for all patients at once: rank the clusters for
    every patient (one batched scoring pass = `rank_cluster_prediction` per patient), take the top cluster, predict
    its mean at the patient's visits, return (assigned cluster, RMSE) per patient.
    XA, YA: NaN-padded patients x visits in the data's own units.  a_draws: the randn each `predict` would draw
    for its new-cluster model (drawn from the global generator, one per patient in order, when omitted)."""
    XA, YA = np.asarray(XA, dtype=np.float64), np.asarray(YA, dtype=np.float64)
    P = XA.shape[0]
    if a_draws is None:
        a_draws = np.array([np.random.randn(1, 1)[0, 0] for _ in range(P)])
    probs = model.predict_batch(XA, YA, a_draws)
    active = np.where(model.allocmodel.Nk > 0)[0]
    # rank_cluster_prediction: argsort(p[active])[::-1][0]
    top = np.array([active[np.argsort(probs[i, active])[::-1][0]] for i in range(P)])
    err = np.full(P, np.nan)
    for k in np.unique(top):
        rows = np.where(top == k)[0]
        keep = ~np.isnan(XA[rows])
        xs = XA[rows][keep]
        mean = model.obsmodel[k].model.predict(xs.reshape(-1, 1))[0].ravel()     # one call per cluster
        off = np.concatenate([[0], np.cumsum(keep.sum(1))])
        for j, r in enumerate(rows):
            y = YA[r][~np.isnan(YA[r])]
            err[r] = calc_error(y, mean[off[j]:off[j + 1]])
    return top, err
