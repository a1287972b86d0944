"""One objective evaluation (inference + gradient) at m = 2 600 for an ncu launch list:
ncu --metrics gpu__time_duration.sum --csv python scripts/one_eval.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi
from mogp_b200.obsmodel import make_spec
os.environ["MOGP_NO_GRAPH"] = "1"
eng = _capi.Engine(0)
spec = make_spec('rbf', True, True, False)
rs = np.random.RandomState(1); m = int(sys.argv[1]) if len(sys.argv) > 1 else 2600
x = np.sort(rs.uniform(0, 8, m)); y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
slot = eng.cluster_create(spec, np.array([0.3, 1.0, 4.0, 0.5]))
eng.cluster_set_data(slot, x, y)
print(eng.cluster_loglik(slot))
t = eng.cluster_optimizer_array(slot)
import time
for rep in range(3):
    t0 = time.time(); f, g = eng.cluster_objective(slot, t + 1e-3 * (rep + 1)); print("eval %.2f ms" % ((time.time() - t0) * 1e3), f)
