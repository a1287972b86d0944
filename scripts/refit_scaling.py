import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi, MoGP_constrained
from mogp_b200.synth import alsfrs_like
from mogp_b200.obsmodel import optimize_many
N, K0 = 10000, 30
eng = _capi.Engine(0)
X, Y, z0 = alsfrs_like(N, seed=0, n_classes=K0, return_labels=True)
z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
mix = MoGP_constrained(X=X, Y=Y, alpha=K0 / np.log10(N), num_iter=3, rand_seed=0, normalize=True, threshold=0.5,
                       noise_variance=0.5, refit='sweep', init_z=z0, engine=eng)
mix.initialize_sampler(K0, True)
models = [mix.obsmodel[k].model for k in mix._active()]
thetas = [eng.cluster_get_theta(m.slot) for m in models]
for blocking in (False, True):
    for nw in (4, 6, 8, 12):
        workers = [_capi.Engine(0, blocking_sync=blocking) for _ in range(nw)]
        ts = []
        for rep in range(4):
            for m, th in zip(models, thetas):
                eng.cluster_set_theta(m.slot, th); m.n_evals = 0
            eng.sync(); t0 = time.time(); optimize_many(models, workers); ts.append(time.time() - t0)
        ne = sum(m.n_evals for m in models)
        print("blocking=%d workers=%2d: %s s (%d evals)" % (blocking, nw, " ".join("%.2f" % t for t in ts), ne), flush=True)
        for w in workers: w.close()
