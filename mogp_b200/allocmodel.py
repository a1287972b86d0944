"""CRP bookkeeping of the DP mixture — host integers, same interface as the reference's
`DPmix` (mogp/allocmodel.py:5-34).  Counts are int64 and bit-exact."""
import numpy as np
from scipy.special import gammaln


class DPmix:
    def __init__(self, alpha):
        self.alpha = alpha
        self.Nk = None
        self.N = None

    def calc_ll(self):                         # allocmodel.py:12-16
        occ = self.Nk[self.Nk > 0]
        return (gammaln(self.alpha) - gammaln(self.N + self.alpha)
                + occ.shape[0] * np.log(self.alpha) + np.sum(gammaln(occ)))

    def calc_suffstats(self, z):               # allocmodel.py:18-20
        self.Nk = np.bincount(z)

    def calc_score(self, curr_k):              # allocmodel.py:22-27
        self.Nk[curr_k] -= 1
        weights = np.hstack([self.Nk, self.alpha])
        return np.log(weights + 1e-16), np.where(weights > 0)[0]

    def update_suffstats(self, new_k):         # allocmodel.py:29-34
        if new_k > self.Nk.shape[0] - 1:
            self.Nk = np.hstack([self.Nk, 1])
        else:
            self.Nk[new_k] += 1
