"""Refit throughput of the 10k-patient / 30-cluster state as a function of the number of lanes of mogp_refit_clusters
(and, through the environment, of CUDA_DEVICE_MAX_CONNECTIONS / MOGP_OPTIMIZER):
    python scripts/refit_lanes.py 4 8 12 16 32
Every measurement starts from the same snapshot (clusters refactorised at the snapshot's hyper-parameters)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mogp_b200 import _capi, MoGP_constrained  # noqa: E402

lanes = [int(v) for v in sys.argv[1:]] or [32]
args = argparse.Namespace(patients=10000, clusters=30)
X, Y, kw = bench.workload(args, 0)
eng = _capi.Engine(0)
mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
mix.initialize_sampler(30, True)
mix.sweep()            # one sweep + refit: a realistic warm state
snap = mix.snapshot()
print("CUDA_DEVICE_MAX_CONNECTIONS =", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "optimizer =",
      os.environ.get("MOGP_OPTIMIZER", "native"))
for L in lanes:
    mix.restore(snap)
    # perturb the hyper-parameters a little so that the refit has real work to do (a restored optimum converges at once)
    _z, Nk, slots = eng.chain_state(10000)
    for k in np.where(Nk > 0)[0]:
        th = eng.cluster_get_theta(int(slots[k]))
        eng.cluster_set_theta(int(slots[k]), th * np.array([1.1, 1.0, 0.9, 1.15]))
    eng.sync()
    mix.refit_lanes = L
    e0 = mix.n_refit_evals
    c0 = eng.refit_counters()
    t0 = time.perf_counter()
    mix.refit_all()
    dt = time.perf_counter() - t0
    c1 = eng.refit_counters()
    ev = mix.n_refit_evals - e0
    dev = c1["launched"] - c0["launched"]
    m = np.array([eng.cluster_size(int(s)) for s in slots if s >= 0], dtype=float)
    flops = dev / len(m) * np.sum(m ** 3 + 12 * m ** 2)
    print("lanes %2d: %.3f s, %d evaluations (%d on the device), %.2f ms per device evaluation, ~%.1f TFLOP/s aggregate"
          % (L, dt, ev, dev, dt / max(dev, 1) * 1e3, flops / dt / 1e12))
