"""CPU restatement of the reference's post-processing / forecasting helpers — TEST INFRASTRUCTURE ONLY.
Follows /root/reference/analysis/model_postprocessing.py:6-104 and analysis/analysis_utils.py:80-92,493-521,
loop for loop, on the oracle's classes."""
import numpy as np

from .reference_model import GPobs, rank_cluster_prediction


def find_cluster_jumps(obs, thresh=10, window=1):                      # model_postprocessing.py:21-32
    min_point = np.min(obs.X)
    while (min_point + window) < np.max(obs.X):
        y_pred_mean, _ = obs.model.predict(np.array([min_point, min_point + window]).reshape(-1, 1))
        if (y_pred_mean[1] - y_pred_mean[0]) > thresh:
            return False
        min_point += 0.1
    return True


def check_model_monotonicity(model, thresh=10, window=1, num_obs=20):   # model_postprocessing.py:6-19
    model_flag = True
    nc = len(np.where(model.allocmodel.Nk > 0)[0])
    idx = np.argsort(-model.allocmodel.Nk)[0:nc]
    if num_obs is not None:
        idx = idx[:num_obs]
    for k in idx:
        model_flag = model_flag and find_cluster_jumps(obs=model.obsmodel[k], thresh=thresh, window=window)
    return model_flag


def generate_clust_obs(model, clust, optimize_params=True):            # model_postprocessing.py:76-104
    xclust, yclust = model.X[clust], model.Y[clust]
    new_obs = GPobs(xclust[~np.isnan(xclust)].reshape(-1, 1), yclust[~np.isnan(yclust)].reshape(-1, 1),
                    kernel=model.kernel, mean_func=model.mean_func, signal_variance=model.signal_variance,
                    signal_variance_fix=model.signal_variance_fix, noise_variance=model.noise_variance,
                    noise_variance_fix=model.noise_variance_fix)
    if optimize_params:
        new_obs.model.optimize()
    model.allocmodel.Nk = np.hstack([model.allocmodel.Nk, xclust.shape[0]])
    new_comp = np.where(model.allocmodel.Nk > 0)[0][-1]
    model.z[clust] = new_comp
    model.obsmodel[new_comp] = new_obs


def split_nonmonotonic_clusters(model, rand_seed=0, thresh=10):        # model_postprocessing.py:48-73
    from sklearn.cluster import KMeans
    nc = len(np.where(model.allocmodel.Nk > 0)[0])
    idx = np.argsort(-model.allocmodel.Nk)[0:nc]
    for k in idx:
        if find_cluster_jumps(obs=model.obsmodel[k], thresh=thresh) is False:
            ind_z = np.where(model.z == k)[0]
            kmeans = KMeans(n_clusters=2, random_state=rand_seed).fit(model.X[ind_z][:, 1].reshape(-1, 1))
            split = np.array(kmeans.labels_)
            model.obsmodel[k] = None
            model.allocmodel.Nk[k] = 0
            generate_clust_obs(model, ind_z[np.where(split == 0)])
            generate_clust_obs(model, ind_z[np.where(split == 1)])
    model._apply_normalizer()


def rank_and_forecast_errors(model, XA, YA):                           # analysis_utils.py:80-92, 493-521
    top, err = [], []
    for i in range(XA.shape[0]):
        cur_x = XA[i][~np.isnan(XA[i])]
        cur_y = YA[i][~np.isnan(YA[i])]
        rank_clust, _ = rank_cluster_prediction(model, cur_x, cur_y)
        k = rank_clust[0]
        y_pred = model.obsmodel[k].model.predict(cur_x.reshape(-1, 1))[0].transpose()
        top.append(k)
        err.append(np.sqrt(((y_pred - cur_y) ** 2).mean()) if len(cur_y) > 0 else np.nan)
    return np.array(top), np.array(err)
