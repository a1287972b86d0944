"""Full-size check of the rank-update path: after one 10k-patient sweep (thousands of row deletions / appends at
fixed theta) every cluster's log marginal and a probe score must equal those of a fresh factorisation."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi, MoGP_constrained
from mogp_b200.synth import alsfrs_like
N, K0 = 10000, 30
eng = _capi.Engine(0)
X, Y, z0 = alsfrs_like(N, seed=0, n_classes=K0, return_labels=True)
z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
mix = MoGP_constrained(X=X, Y=Y, alpha=K0 / np.log10(N), num_iter=3, rand_seed=0, normalize=True, threshold=0.5,
                       noise_variance=0.5, refit='sweep', init_z=z0, engine=eng)
mix.initialize_sampler(K0, True)
mix._sync_state(); mix.refit_all()
perm, a, u = mix._draws_for_sweep(N)
t0 = time.time(); chosen, mm = eng.chain_sweep(perm, a, u); print("sweep %.2fs, min margin %.2e" % (time.time() - t0, mm))
mix._sync_state()
moved = int(np.sum(mix.z != z0)); print("moved", moved)
rs = np.random.RandomState(1)
xq = np.sort(np.r_[0.0, rs.uniform(0.3, 8, 8)]); yq = rs.randn(9)
worst_lml = worst_sc = 0.0
for k in mix._active():
    s = mix.obsmodel[k].model.slot
    lml_u = eng.cluster_loglik(s)
    sc_u = eng.score_pairs(np.array([0, 9]), xq, yq, [s])[0, 0]
    x, y = eng.cluster_get_data(s)
    f = eng.cluster_create(mix._spec(), eng.cluster_get_theta(s))
    eng.cluster_set_data(f, x, y)
    lml_f = eng.cluster_loglik(f)
    sc_f = eng.score_pairs(np.array([0, 9]), xq, yq, [f])[0, 0]
    eng.cluster_free(f)
    worst_lml = max(worst_lml, abs(lml_u - lml_f) / abs(lml_f)); worst_sc = max(worst_sc, abs(sc_u - sc_f) / abs(sc_f))
print("after one sweep of rank updates: worst relative LML difference %.2e, worst relative score difference %.2e (vs fresh factorisation)" % (worst_lml, worst_sc))
