"""Per-operation timings at the headline configuration (10k patients, 30 clusters)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi, MoGP_constrained
from mogp_b200.synth import alsfrs_like, block_init

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
K0 = int(sys.argv[2]) if len(sys.argv) > 2 else 30
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
eng = _capi.Engine(0)
t0 = time.time(); X, Y, z0 = alsfrs_like(N, seed=0, n_classes=K0, return_labels=True); print("gen %.2fs" % (time.time() - t0))
z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
mix = MoGP_constrained(X=X, Y=Y, alpha=K0 / np.log10(N), num_iter=3, rand_seed=0, normalize=True, threshold=0.5,
                       noise_variance=0.5, refit='sweep', init_z=z0, engine=eng)
t0 = time.time(); mix.initialize_sampler(K0, True); eng.sync(); print("init (30 inferences) %.3fs" % (time.time() - t0))
_z, Nk, slots = eng.chain_state(N)
s0 = int(slots[0]); print("m of cluster 0:", eng.cluster_size(s0))
for rep in range(3):
    t0 = time.time(); eng.cluster_set_theta(s0, eng.cluster_get_theta(s0)); t1 = time.time()
    g = eng.cluster_lml_grad(s0); t2 = time.time()
    print("infer %.2f ms   grad %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
perm, a, u = mix._draws_for_sweep(N)
eng.profile(True)
for chunk in range(3):
    sl = slice(chunk * steps, (chunk + 1) * steps)
    l0 = eng.launch_count()
    t0 = time.time(); chosen, mm = eng.chain_sweep(perm[sl], a[sl], u[sl]); eng.sync(); dt = time.time() - t0
    moved = int(np.sum(chosen != _z[perm[sl]]))
    pr = eng.profile_read()
    print("chunk %d: %.1f us/step, moved %d/%d, launches/step %.1f, score_gemm %.1f us/launch, alg %.1f GB/s %.2f TFLOP/s"
          % (chunk, dt / steps * 1e6, moved, steps, (eng.launch_count() - l0) / steps, pr['ms'] / max(pr['launches'], 1) * 1e3,
             pr['alg_bytes'] / (pr['ms'] * 1e-3) / 1e9, pr['alg_flops'] / (pr['ms'] * 1e-3) / 1e12))
    _z, Nk, slots = eng.chain_state(N)
t0 = time.time(); mix._sync_state(); mix.obsmodel[0].model.optimize(); dt = time.time() - t0
print("one refit: %.2fs, %d evals" % (dt, mix.obsmodel[0].model.n_evals))
# concurrent refit scaling
from mogp_b200.obsmodel import optimize_many
import threading
models = [mix.obsmodel[k].model for k in mix._active()]
thetas = [eng.cluster_get_theta(m.slot) for m in models]
for nw in (0, 2, 4, 8, 16):
    for m, th in zip(models, thetas):
        eng.cluster_set_theta(m.slot, th)
        m.n_evals = 0
    workers = [_capi.Engine(0) for _ in range(nw)]
    eng.sync()
    t0 = time.time(); optimize_many(models, workers); dt = time.time() - t0
    ne = sum(m.n_evals for m in models)
    print("refit all %d clusters, %2d workers: %.2f s, %d evals, %.2f ms/eval" % (len(models), nw, dt, ne, dt / ne * 1e3))
# raw concurrency of the objective without scipy: each thread evaluates its cluster 10 times
def raw(nw):
    workers = [_capi.Engine(0) for _ in range(nw)]
    ts = [eng.cluster_optimizer_array(m.slot) for m in models]
    def run(i):
        for rep in range(10):
            eng.cluster_objective(models[i].slot, ts[i], worker=workers[i % nw])
    eng.sync(); t0 = time.time()
    th = [threading.Thread(target=lambda lst=list(range(w, len(models), nw)): [run(i) for i in lst]) for w in range(nw)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.time() - t0
    print("raw objective x10 on %d clusters, %2d threads: %.2f s -> %.2f ms/eval" % (len(models), nw, dt, dt / (10 * len(models)) * 1e3))
for nw in (1, 4, 16):
    raw(nw)
