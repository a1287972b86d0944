"""Generate tests/golden/likelihood_golden.npz from the CPU oracle (run here, in the build container).

The reference itself cannot be imported (GPy absent), so these vectors come from the oracle
AFTER it has been pinned to the reference's executed tutorial notebook (tests/test_oracle_kat.py); they freeze
the oracle's numbers for the paths the notebook does not print (linear kernel, mean_func=False, un-fixed signal
variance, predict, new-cluster marginal) so that a later edit of the oracle cannot move them silently, and give
the GPU tests a fixture that does not depend on the installed NumPy/SciPy.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.gpy_restated import GPRegressionOracle  # noqa: E402

CASES = [  # kernel, mean_func, sig_fix, noise_fix, theta (A, k0, k1, noise), m
    ('rbf', True, True, False, (0.28, 1.0, 1.58, 0.0133), 37),
    ('rbf', True, False, False, (0.1, 0.7, 2.5, 0.2), 65),
    ('rbf', False, True, False, (0.0, 1.0, 1.0, 0.05), 130),
    ('linear', True, True, False, (0.4, 1.0, 1.0, 0.5), 70),
    ('linear', False, False, False, (0.0, 0.6, 0.8, 0.1), 33),
]


def make_case(i, case):
    kernel, mf, sfix, nfix, th, m = case
    rs = np.random.RandomState(1000 + i)
    x = np.sort(rs.uniform(0, 8, m))
    x[: max(1, m // 8)] = 0.0
    y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
    o = GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name=kernel, signal_variance=th[1],
                           signal_variance_fix=sfix, noise_variance=th[3], noise_variance_fix=nfix, mean_func=mf,
                           lengthscale=th[2], a_draw=th[0] if mf else None)
    if kernel == 'linear':
        o.param_array[o._off + 1] = th[2]
        o._inference()
    xs = rs.uniform(0, 9, 9)
    ys = rs.randn(9)
    mu, var = o.predict(xs.reshape(-1, 1))
    lpd = o.log_predictive_density(xs.reshape(-1, 1), ys.reshape(-1, 1)).ravel()
    grad = np.zeros(4)
    grad[(0 if mf else 1):] = o._grad_lml
    lml = o.log_likelihood()
    t = o.optimizer_array() + rs.uniform(-0.3, 0.3, o.optimizer_array().size)
    f, g = o.objective_grads(t)
    # new-cluster marginal of the 9 query points as one patient (GPobs.__init__, obsmodel.py:56), A = |a|
    a_draw = rs.randn()
    nc = GPRegressionOracle(np.sort(xs).reshape(-1, 1), ys.reshape(-1, 1), kernel_name=kernel, signal_variance=1.0,
                            noise_variance=0.5, mean_func=mf, a_draw=a_draw if mf else None)
    return dict(x=x, y=y, xs=xs, ys=ys, theta=np.array(th), lml=lml, grad=grad, mu=mu.ravel(), var=var.ravel(), lpd=lpd,
                t=t, f=f, g=g, a_draw=a_draw, new_ll=nc.log_likelihood(),
                spec=np.array([0 if kernel == 'rbf' else 1, int(mf), int(sfix), int(nfix)]))


if __name__ == "__main__":
    out = {}
    for i, case in enumerate(CASES):
        for k, v in make_case(i, case).items():
            out["c{}_{}".format(i, k)] = np.asarray(v)
    out["n_cases"] = np.array(len(CASES))
    np.savez(os.path.join(HERE, "likelihood_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "likelihood_golden.npz"))
