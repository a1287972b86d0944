"""Where does the concurrent end-of-sweep refit spend its wall time: inside the engine (GPU + waits) or in Python/SciPy?"""
import sys, os, time, threading
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
mix.sweep()
inside = {}
orig = eng.lib.mogp_cluster_objective_on
def timed(*a):
    t0 = time.perf_counter(); r = orig(*a); dt = time.perf_counter() - t0
    k = threading.get_ident(); inside[k] = inside.get(k, 0.0) + dt
    return r
for nw in (int(a) for a in (sys.argv[1:] or ["12"])):
    mix.refit_concurrency = nw
    mix._workers = None
    perm, a, u = mix._draws_for_sweep(N)
    eng.chain_sweep(perm, a, u); mix._sync_state()
    eng.lib.mogp_cluster_objective_on = timed
    inside.clear(); mix.n_refit_evals = 0
    t0 = time.perf_counter(); mix.refit_all(); wall = time.perf_counter() - t0
    eng.lib.mogp_cluster_objective_on = orig
    tot = sum(inside.values())
    print("%2d threads: wall %.3f s, %d evals, sum of time inside the engine call %.3f s = %.2f x wall (%.0f %% of thread time), %.2f ms per call"
          % (nw, wall, mix.n_refit_evals, tot, tot / wall, 100 * tot / (wall * max(len(inside), 1)), 1e3 * tot / max(mix.n_refit_evals, 1)))
