"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck), sizes kept tiny."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi, MoGP_constrained
from mogp_b200.obsmodel import make_spec
from mogp_b200.synth import alsfrs_like, block_init
eng = _capi.Engine(0)
rs = np.random.RandomState(0)
spec = make_spec('rbf', True, True, False)
m = 150
x = np.sort(rs.uniform(0, 8, m)); y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
s = eng.cluster_create(spec, np.array([0.3, 1.0, 2.0, 0.1]))
eng.cluster_set_data(s, x, y)
print("lml", eng.cluster_loglik(s), "grad", eng.cluster_lml_grad(s))
t = eng.cluster_optimizer_array(s)
print("objective (graph)", eng.cluster_objective(s, t + 0.01))
print("objective (graph again)", eng.cluster_objective(s, t - 0.01))
print("leave-out", eng.cluster_leave_out_score(s, 60, 9, 1))
eng.cluster_remove_rows(s, 60, 9)
eng.cluster_append(s, x[60:69], y[60:69])
print("lml after rm+append", eng.cluster_loglik(s))
print("predict", eng.cluster_predict(s, np.array([0.5, 3.0, 7.0])))
X, Y = alsfrs_like(24, seed=1)
mix = MoGP_constrained(X=X, Y=Y, alpha=1.0, num_iter=2, rand_seed=0, normalize=True, threshold=0.5, noise_variance=0.5,
                       refit='sweep', init_z=block_init(24, 2, X), engine=eng)
mix.sample()
print("chain ok", mix.allocmodel.Nk)
print("predict_batch", mix.predict_batch(X[:3], Y[:3] * mix.std + mix.mean, np.array([0.1, -0.3, 0.7])).sum(1))
eng.close()
