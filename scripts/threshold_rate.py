"""How many (patient, cluster) pairs does the monotonicity rule (mogp_constrained.py:254-263) exclude at the headline
configuration?  For those pairs only the predictive MEAN at the threshold visit is needed - the O(m^2) variance product
is not (DESIGN.md 'next')."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi, MoGP_constrained
from mogp_b200.synth import alsfrs_like
from mogp_b200.utils import pack_patients

N, K0 = 10000, 30
eng = _capi.Engine(0)
X, Y, z0 = alsfrs_like(N, seed=0, n_classes=K0, return_labels=True)
z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
mix = MoGP_constrained(X=X, Y=Y, alpha=K0 / np.log10(N), num_iter=3, rand_seed=0, normalize=True, threshold=0.5,
                       noise_variance=0.5, refit='sweep', init_z=z0, engine=eng)
mix.initialize_sampler(K0, True)
for sweep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    off, x, y = pack_patients(mix.X, mix.Y)
    _z, Nk, slots = eng.chain_state(N)
    act = np.array([s for s in slots if s >= 0], dtype=np.int32)
    sel = np.arange(0, N, 5)
    qoff = np.zeros(len(sel) + 1, dtype=np.int64)
    qx, qy = [], []
    for i, n in enumerate(sel):
        qx.append(x[off[n]:off[n + 1]]); qy.append(y[off[n]:off[n + 1]]); qoff[i + 1] = qoff[i] + (off[n + 1] - off[n])
    score, mu = eng.score_pairs(qoff, np.concatenate(qx), np.concatenate(qy), act, thr_index=1, want_mu=True)
    ythr = mix.Y[sel, 1][:, None]
    excl = (ythr - mu) > 0.5
    p = np.exp(score - score.max(axis=1, keepdims=True)); p[excl] = 0; p /= p.sum(axis=1, keepdims=True)
    print("before sweep %d: %d clusters, %.1f %% of (patient, cluster) pairs excluded by the threshold; pairs with p > 1e-12 among the rest: %.1f %%"
          % (sweep, len(act), 100 * excl.mean(), 100 * (p[~excl] > 1e-12).mean()))
    mix.sweep()
