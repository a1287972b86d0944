"""Generate the chain-level golden fixtures from the CPU oracle (run here, in the build container; minutes to
half an hour of CPU).  The oracle is pinned to the reference's executed tutorial notebook
(tests/test_oracle_kat.py) and, through tests/test_reference_sampler.py, its sampler restatement is checked against the
reference's own unmodified sampler code; these fixtures freeze what it computes at BASELINE.json's configs[0] and
configs[1] shapes so that the GPU tests can compare the engine's free-running chains with it without paying for
the oracle on the GPU box.

    python tests/golden/make_chain_golden.py cfg1     # MoGP_constrained, 200 patients, 100 iterations, reference schedule
    python tests/golden/make_chain_golden.py cfg2     # MoGP (unconstrained), 2 000 patients, refit='sweep', 2 sweeps

k-means labels are stored with the fixture: sklearn numbers the same partition differently between versions
(SURVEY.md section 8c), and k-means is outside the likelihood path.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from mogp_b200.synth import alsfrs_like  # noqa: E402  (plain NumPy generator, no engine)


def kmeans_init(X, K, seed):
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=K, random_state=seed, n_init=10).fit(X[:, 1].reshape(-1, 1))
    return km.labels_.astype(np.int32)


def pad(rows):
    w = max(len(r) for r in rows)
    out = np.full((len(rows), w), -1, dtype=np.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def run(name):
    from threadpoolctl import threadpool_limits
    if name == 'cfg1':
        N, iters, K0 = 200, 100, 4
        X, Y = alsfrs_like(N, seed=0)
        z0 = kmeans_init(X, K0, 0)
        kw = dict(alpha=K0 / np.log10(N), num_iter=iters, rand_seed=0, normalize=True, threshold=0.5, noise_variance=0.5,
                  init_z=z0)
        cls = oracle.MoGP_constrained
    elif name == 'cfg2':
        N, iters, K0 = 2000, 3, 40
        X, Y = alsfrs_like(N, seed=1)
        z0 = kmeans_init(X, K0, 0)
        kw = dict(alpha=K0 / np.log10(N), num_iter=iters, rand_seed=0, normalize=True, noise_variance=0.5, refit='sweep',
                  init_z=z0)
        cls = oracle.MoGP
    else:
        raise SystemExit("cfg1 | cfg2")
    orc = cls(X=X.copy(), Y=Y.copy(), **kw)
    orc.track_margin = True
    t0 = time.time()
    with threadpool_limits(1 if name == 'cfg1' else 4):
        orc.sample()
    dt = time.time() - t0
    act = np.where(orc.allocmodel.Nk > 0)[0]
    theta = np.array([np.asarray(orc.obsmodel[k].model.param_array, dtype=float) for k in act])
    out = os.path.join(HERE, "chain_{}.npz".format(name))
    np.savez_compressed(out, z0=z0, z=np.asarray(orc.z, dtype=np.int64), Nk=np.asarray(orc.allocmodel.Nk, dtype=np.int64),
                        trace=pad([list(t) for t in orc.trace]), lobs=orc.lobs, lalloc=orc.lalloc,
                        active=act, theta=theta, min_margin=orc.min_margin, seconds=dt)
    print(name, "done in", round(dt), "s; occupancy", list(orc.allocmodel.Nk[act]), "min margin", orc.min_margin, "->", out)


if __name__ == "__main__":
    run(sys.argv[1])
