"""Synthetic sparse ALSFRS-R-shaped trajectories for tests and bench.py.

Per patient: n_visits ~ U{3..10}; first visit t0 ~ U(0.2, 3) years after onset, later visits spread so that
all t are in (0, 8]; a trajectory family (linear decline / early sigmoid / late sigmoid) with a per-patient
rate; y = clip(round(f(t) + N(0, 1.5^2)), 0, 48); onset anchor (x=0, y=48) prepended
(analysis/process_data_for_mogp_experiments.py:33-36); NaN-padded to T = 11 columns.
"""
import numpy as np


def alsfrs_like(n_patients, seed=0, max_visits=10, min_visits=3, anchor=True, n_classes=None, return_labels=False):
    """n_classes=None: every patient draws its own (family, rate).  n_classes=C: patients belong to C latent
    progression classes (family x rate x onset on a grid, small per-patient jitter) — the shape of a sampler
    that has found its clusters, which is the state BASELINE.json's 10k-patient configuration names
    (~30 active clusters)."""
    rs = np.random.RandomState(seed)
    labels = np.zeros(n_patients, dtype=np.int32)
    T = max_visits + (1 if anchor else 0)
    X = np.full((n_patients, T), np.nan)
    Y = np.full((n_patients, T), np.nan)
    for i in range(n_patients):
        nv = rs.randint(min_visits, max_visits + 1)
        t0 = rs.uniform(0.2, 3.0)
        gaps = rs.uniform(0.1, 1.0, nv - 1)
        t = np.concatenate([[t0], t0 + np.cumsum(gaps)])
        if t[-1] > 8.0:
            t = t0 + (t - t0) * (8.0 - t0) / (t[-1] - t0)
        if n_classes is None:
            fam = rs.randint(3)
            rate = rs.uniform(0.5, 1.5)
            mid = 1.5 if fam == 1 else 4.0
        else:
            c = rs.randint(n_classes)
            labels[i] = c
            fam = c % 3
            q = (c // 3) / max(1, (n_classes + 2) // 3 - 1) if n_classes > 3 else 0.5
            rate = (0.3 + 1.5 * q) * rs.uniform(0.97, 1.03)
            mid = (1.0 + 2.0 * q) if fam == 1 else (3.0 + 3.0 * q)
        if fam == 0:
            f = 48.0 - 6.0 * rate * t
        elif fam == 1:
            f = 48.0 / (1.0 + np.exp(2.0 * rate * (t - mid)))
        else:
            f = 48.0 / (1.0 + np.exp(1.5 * rate * (t - mid)))
        y = np.clip(np.round(f + rs.randn(nv) * 1.5), 0, 48)
        o = 1 if anchor else 0
        if anchor:
            X[i, 0], Y[i, 0] = 0.0, 48.0
        X[i, o:o + nv] = t
        Y[i, o:o + nv] = y
    if return_labels:
        return X, Y, labels
    return X, Y


def block_init(n_patients, n_clusters, X=None):
    """A deterministic initial partition with n_clusters roughly equal clusters: patients ranked by their
    first real visit time (the quantity the reference's k-means init clusters on) and cut into quantile blocks."""
    if X is None:
        return (np.arange(n_patients) * n_clusters // n_patients).astype(np.int32)
    col = 1 if X.shape[1] > 1 else 0
    order = np.argsort(X[:, col], kind='stable')
    z = np.empty(n_patients, dtype=np.int32)
    z[order] = (np.arange(n_patients) * n_clusters // n_patients).astype(np.int32)
    return z


def alsfrs_like_fast(n_patients, seed=0, max_visits=10, min_visits=3, n_classes=30):
    """Vectorised variant of `alsfrs_like` for very large query sets (batch prediction, 1M patients):
    same shape (onset anchor + 3..10 visits in (0, 8], NaN-padded to 11 columns), class-structured."""
    rs = np.random.RandomState(seed)
    T = max_visits + 1
    nv = rs.randint(min_visits, max_visits + 1, n_patients)
    t0 = rs.uniform(0.2, 3.0, n_patients)
    gaps = rs.uniform(0.1, 1.0, (n_patients, max_visits))
    gaps[:, 0] = 0.0
    t = t0[:, None] + np.cumsum(gaps, axis=1)
    last = t[np.arange(n_patients), nv - 1]
    scale = np.where(last > 8.0, (8.0 - t0) / np.maximum(last - t0, 1e-12), 1.0)
    t = t0[:, None] + (t - t0[:, None]) * scale[:, None]
    c = rs.randint(n_classes, size=n_patients)
    fam = c % 3
    q = (c // 3) / max(1, (n_classes + 2) // 3 - 1) if n_classes > 3 else np.full(n_patients, 0.5)
    rate = (0.3 + 1.5 * q) * rs.uniform(0.97, 1.03, n_patients)
    mid = np.where(fam == 1, 1.0 + 2.0 * q, 3.0 + 3.0 * q)
    lin = 48.0 - 6.0 * rate[:, None] * t
    sig1 = 48.0 / (1.0 + np.exp(2.0 * rate[:, None] * (t - mid[:, None])))
    sig2 = 48.0 / (1.0 + np.exp(1.5 * rate[:, None] * (t - mid[:, None])))
    f = np.where(fam[:, None] == 0, lin, np.where(fam[:, None] == 1, sig1, sig2))
    y = np.clip(np.round(f + rs.randn(n_patients, max_visits) * 1.5), 0, 48)
    mask = np.arange(max_visits)[None, :] < nv[:, None]
    X = np.full((n_patients, T), np.nan)
    Y = np.full((n_patients, T), np.nan)
    X[:, 0], Y[:, 0] = 0.0, 48.0
    X[:, 1:] = np.where(mask, t, np.nan)
    Y[:, 1:] = np.where(mask, y, np.nan)
    return X, Y
