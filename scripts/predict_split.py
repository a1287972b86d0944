import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.special import logsumexp
from mogp_b200 import _capi, MoGP_constrained
from mogp_b200.synth import alsfrs_like, alsfrs_like_fast
from mogp_b200.utils import pack_patients
from mogp_b200.obsmodel import initial_theta
K = 40
eng = _capi.Engine(0)
Xt, Yt, z0 = alsfrs_like(3000, seed=0, n_classes=K, return_labels=True)
z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
mix = MoGP_constrained(X=Xt, Y=Yt, alpha=K / np.log10(3000), rand_seed=0, normalize=True, noise_variance=0.5, init_z=z0, engine=eng)
mix.initialize_sampler(K, True)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
Xq, Yq = alsfrs_like_fast(P, seed=2, n_classes=K)
a = np.random.RandomState(3).randn(P)
mix.predict_batch(Xq[:5000], Yq[:5000], a[:5000])
for rep in range(2):
    t0 = time.time(); Yn = (Yq - mix.mean) / mix.std; off, x, y = pack_patients(Xq, Yn); t1 = time.time()
    active = mix._active(); slots = np.array([mix.obsmodel[k].model.slot for k in active], dtype=np.int32)
    eng.profile(True); eng.profile_read()
    sc = eng.score_pairs(off, x, y, slots); t2 = time.time()
    pr = eng.profile_read(); eng.profile(False)
    th0 = initial_theta('rbf', 1.0, 0.5, 4., 0.0)
    nl = eng.new_cluster_ll(mix._spec(mean_func=True), th0, off, x, y, a); t3 = time.time()
    score = np.tile(np.log(np.hstack([mix.allocmodel.Nk, mix.allocmodel.alpha]) + 1e-16), (P, 1))
    score[:, active] += sc; score[:, -1] += nl
    p = np.exp(score - logsumexp(score, axis=1, keepdims=True)); t4 = time.time()
    print("pack %.3f  score_pairs %.3f (gemm %.3f s, %.1f TF)  new_ll %.3f  softmax %.3f  total %.3f" % (t1-t0, t2-t1, pr['ms']*1e-3, pr['alg_flops']/(pr['ms']*1e-3)/1e12, t3-t2, t4-t3, t4-t0))
