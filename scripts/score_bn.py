"""Scoring-pass time by query width at the headline configuration (10k patients, 30 clusters of m ~ 2600):
R query columns (R/8 patients x 8 visits) against every cluster, k_score_gemm<BN> timed with CUDA events."""
import sys, os
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
_z, Nk, slots = eng.chain_state(N)
slots = np.array([s for s in slots if s >= 0], dtype=np.int32)
ms = [eng.cluster_size(int(s)) for s in slots]
print("clusters", len(slots), "m mean", np.mean(ms), "sum m^2/2 * 8 B =", sum(m * m / 2 * 8 for m in ms) / 1e9, "GB")
rs = np.random.RandomState(0)
for R in [int(a) for a in sys.argv[1:]] or [8, 16, 24, 32, 40, 48, 64]:
    P = max(1, R // 8)
    off = np.arange(P + 1, dtype=np.int64) * (R // P)
    x = np.sort(rs.uniform(0, 8, R)); y = rs.randn(R)
    eng.profile(True); eng.profile_read()
    for rep in range(12):
        eng.score_pairs(off, x, y, slots)
        if rep == 1:
            eng.profile_read()
    pr = eng.profile_read(); eng.profile(False)
    us = pr['ms'] / pr['launches'] * 1e3
    print("R=%2d: %.1f us/launch  %.2f TFLOP/s  %.0f GB/s (factor once)  %.1f us per 8 columns"
          % (R, us, pr['alg_flops'] / (pr['ms'] * 1e-3) / 1e12, pr['alg_bytes'] / (pr['ms'] * 1e-3) / 1e9, us / (R / 8)))
