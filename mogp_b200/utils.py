"""Helpers of the reference's `mogp/utils.py` that sit on the likelihood path."""
import numpy as np


def rank_cluster_prediction(model, x_new, y_new):                  # utils.py:9-29
    """Rank the active clusters of `model` for one new patient (argsort of `predict`)."""
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


def pack_patients(X, Y):
    """NaN-padded N x T matrices -> CSR (off, x, y) of the non-NaN cells, row-major
    (the compaction the reference does per row at mogp_constrained.py:246-247)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    keep_x, keep_y = ~np.isnan(X), ~np.isnan(Y)
    nx, ny = keep_x.sum(1), keep_y.sum(1)
    if np.any(nx != ny):
        raise ValueError("X and Y have different numbers of non-NaN cells in some row")
    off = np.zeros(X.shape[0] + 1, dtype=np.int64)
    np.cumsum(nx, out=off[1:])
    return off, X[keep_x], Y[keep_y]


def plot_mogp(*a, **k):                                            # utils.py:50-79
    raise NotImplementedError("plotting (matplotlib through GPy) is outside the likelihood engine")
