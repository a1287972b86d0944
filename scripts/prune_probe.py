"""The scoring product of one look-ahead batch on its own, with and without the device-side monotonicity pruning:
    python scripts/prune_probe.py            # unpruned, 48 / 56 / 64 columns
    MOGP_PROBE_PRUNE=4 python scripts/prune_probe.py   # every 4th patient excluded under every cluster
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mogp_b200 import _capi, MoGP_constrained  # noqa: E402
from mogp_b200.utils import pack_patients  # noqa: E402

args = argparse.Namespace(patients=10000, clusters=30)
X, Y, kw = bench.workload(args, 0)
eng = _capi.Engine(0)
mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
mix.initialize_sampler(30, True)
off, px, py = pack_patients(mix.X, mix.Y)
_z, _Nk, slots = eng.chain_state(10000)
act = np.array([s for s in slots if s >= 0], dtype=np.int32)
for cols in (48, 56, 64):
    take, tot = [], 0
    for n in np.random.RandomState(1).permutation(10000):
        ni = int(off[n + 1] - off[n])
        if tot + ni > cols:
            continue
        take.append(n); tot += ni
        if tot >= cols - 2:
            break
    qoff = np.zeros(len(take) + 1, dtype=np.int64)
    qx, qy = [], []
    for i, n in enumerate(take):
        qx.append(px[off[n]:off[n + 1]]); qy.append(py[off[n]:off[n + 1]])
        qoff[i + 1] = qoff[i] + (off[n + 1] - off[n])
    qx, qy = np.concatenate(qx), np.concatenate(qy)
    eng.profile(True); eng.profile_read(reset=True)
    for rep in range(12):
        sc = eng.score_pairs(qoff, qx, qy, act, thr_index=1)
        if rep == 1:
            eng.profile_read(reset=True)
    pi = eng.profile_read(reset=True); eng.profile(False)
    print("probe=%s columns %d patients %d: %.1f us per launch, %d pairs -inf of %d" % (
        os.environ.get("MOGP_PROBE_PRUNE", "0"), tot, len(take), pi["ms"] / max(pi["launches"], 1) * 1e3,
        int(np.sum(np.isinf(sc))), sc.size))
