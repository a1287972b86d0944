"""CPU restatement of the rank-n SCORE identities the engine's look-ahead pipeline uses (mogp_b200/csrc/update.cuh:
k_alg_leave_*, k_alg_join_*) - TEST INFRASTRUCTURE ONLY.

The reference has no such step: after a patient moves it refactorises the two clusters (`GPobs.update_data` /
`update_suffstats`, mogp/obsmodel.py:68-73, 87-94) and the next patient is scored from the new factors
(`calc_score`, obsmodel.py:75-81).  The engine instead corrects the scores it has ALREADY computed for the patients
ahead, from T_q = M k_q (M = chol(Ky)^-1), s_q = |T_q|^2 and mu_q.  These functions state that algebra in NumPy so that
`tests/test_oracle_properties.py` can check it against a refactorised `GPRegressionOracle` without a GPU.
"""
import numpy as np


def leave_block(M, alpha, T_q, s_q, mu_q, rows):
    """Block B = `rows` leaves the cluster.  W = M^T M = Ky^-1, u = (M^T T_q)_B = (W k_q)_B, G = W_BB:
       s' = s - u^T G^-1 u,  mu' = mu - u^T G^-1 alpha_B   (block elimination of Ky^-1)."""
    B = np.asarray(rows)
    b = M[:, B]                          # the block's columns of the inverse factor (zero above the block)
    G = b.T @ b
    u = b.T @ T_q                        # n x (query columns)
    Gi_u = np.linalg.solve(G, u)
    return s_q - np.sum(u * Gi_u, axis=0), mu_q - Gi_u.T @ alpha[B]


def join_block(T_q, s_q, mu_q, T_B, K_Bq, S_BB, r_B):
    """Block B joins the cluster.  S_BB = K_BB + (noise + jitter) I - T_B^T T_B (the block's predictive covariance),
       Ls = chol(S_BB), z_B = Ls^-1 r_B with r_B = y_B - predictive mean at x_B, t = Ls^-1 (K_Bq - T_B^T T_q):
       s' = s + |t|^2,  mu' = mu + t . z_B   (the rows the append adds to T_q)."""
    Ls = np.linalg.cholesky(S_BB)
    t = np.linalg.solve(Ls, K_Bq - T_B.T @ T_q)
    z_B = np.linalg.solve(Ls, r_B)
    return s_q + np.sum(t * t, axis=0), mu_q + t.T @ z_B


def alpha_after_leave(M, alpha, rows):
    """alpha = Ky^-1 r of the cluster without block B, from the old factor (k_rm_alpha):
       alpha'_R = alpha_R - W_RB W_BB^-1 alpha_B with W = M^T M."""
    B = np.asarray(rows)
    keep = np.setdiff1d(np.arange(M.shape[0]), B)
    b = M[:, B]
    W_B = b.T @ M                         # W_{B,:}
    g = np.linalg.solve(W_B[:, B], alpha[B])
    return alpha[keep] - W_B[:, keep].T @ g


def alpha_after_join(M, alpha, T_B, S_BB, r_B):
    """alpha of the cluster with block B appended (k_append_finalize): new rows of the inverse factor
       R = -Ls^-1 T_B^T M, z_B = Ls^-1 r_B:  alpha' = [alpha + R^T z_B ; Ls^-T z_B]."""
    Ls = np.linalg.cholesky(S_BB)
    z_B = np.linalg.solve(Ls, r_B)
    R = -np.linalg.solve(Ls, T_B.T @ M)
    return np.concatenate([alpha + R.T @ z_B, np.linalg.solve(Ls.T, z_B)])
