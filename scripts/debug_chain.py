import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, warnings
import oracle
from mogp_b200 import _capi, MoGP_constrained
from test_edge_cases import ragged_data
alpha = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
X, Y = ragged_data(1)
z0 = (np.arange(40) % 3).astype(np.int32)
kw = dict(alpha=alpha, num_iter=3, rand_seed=2, normalize=True, threshold=0.5, noise_variance=0.5, refit='sweep', init_z=z0)
eng = _capi.Engine(0)
orc = oracle.MoGP_constrained(X=X.copy(), Y=Y.copy(), **kw)
orc.initialize_sampler(orc.num_init_clusters, True); orc.allocmodel.calc_suffstats(orc.z); orc.step_log = []
mix = MoGP_constrained(X=X.copy(), Y=Y.copy(), engine=eng, **kw)
mix.initialize_sampler(mix.num_init_clusters, True)
for sweep in range(2):
    perm = np.random.permutation(np.arange(40))
    for n in perm:
        st = np.random.get_state(); a = np.random.randn(1, 1)[0, 0]; u = np.random.random_sample(); np.random.set_state(st)
        orc.gibbs_step(n); rec = orc.step_log[-1]
        eng.chain_step_begin(n)
        ids, score = eng.chain_step_score(a)
        idx, p, margin = _capi.host_choice(score, u)
        ok = list(ids) == list(rec['active']) and np.max(np.abs(p - rec['p'])) < 1e-9
        if not ok or ids[idx] != rec['z']:
            print("MISMATCH sweep", sweep, "patient", n, "n_i", int((~np.isnan(X[n])).sum()))
            print(" ids", ids, rec['active'])
            print(" engine score", score)
            print(" oracle score", rec['score'])
            print(" p", p, rec['p'], "u", u, "chosen", ids[idx], rec['z'])
            sys.exit(0)
        eng.chain_step_commit(rec['z'], a)
    orc.end_of_sweep_refit()
    mix._sync_state(); mix.refit_all()
    for k in np.where(orc.allocmodel.Nk > 0)[0]:
        print("sweep", sweep, "cluster", k, "theta engine", mix.obsmodel[k].model.param_array, "oracle", orc.obsmodel[k].model.param_array)
print("all steps agree")
