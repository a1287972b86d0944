"""Committed golden vectors (tests/golden/likelihood_golden.npz, made by tests/golden/make_golden.py from the
notebook-pinned oracle): the oracle must still reproduce them (CPU), and the CUDA engine must match them (GPU)."""
import os

import numpy as np
import pytest

from oracle.gpy_restated import GPRegressionOracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "likelihood_golden.npz"))
N = int(G["n_cases"])


def case(i):
    return {k[len("c%d_" % i):]: G[k] for k in G.files if k.startswith("c%d_" % i)}


@pytest.mark.parametrize("i", range(N))
def test_oracle_reproduces_golden(i):
    c = case(i)
    kernel = 'rbf' if c['spec'][0] == 0 else 'linear'
    mf, sfix, nfix = bool(c['spec'][1]), bool(c['spec'][2]), bool(c['spec'][3])
    th = c['theta']
    o = GPRegressionOracle(c['x'].reshape(-1, 1), c['y'].reshape(-1, 1), kernel_name=kernel, signal_variance=th[1],
                           signal_variance_fix=sfix, noise_variance=th[3], noise_variance_fix=nfix, mean_func=mf,
                           lengthscale=th[2], a_draw=th[0] if mf else None)
    if kernel == 'linear':
        o.param_array[o._off + 1] = th[2]
        o._inference()
    assert o.log_likelihood() == pytest.approx(float(c['lml']), rel=1e-11)
    mu, var = o.predict(c['xs'].reshape(-1, 1))
    assert np.allclose(mu.ravel(), c['mu'], rtol=1e-9, atol=1e-12)
    assert np.allclose(var.ravel(), c['var'], rtol=1e-10)
    f, g = o.objective_grads(c['t'])
    assert f == pytest.approx(float(c['f']), rel=1e-11)
    assert np.allclose(g, c['g'], rtol=1e-8, atol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(N))
def test_engine_matches_golden(i):
    from mogp_b200 import _capi
    c = case(i)
    eng = _capi.Engine(0)
    spec = _capi.ModelSpec(*[int(v) for v in c['spec']])
    slot = eng.cluster_create(spec, c['theta'])
    eng.cluster_set_data(slot, c['x'], c['y'])
    assert eng.cluster_loglik(slot) == pytest.approx(float(c['lml']), rel=1e-9)
    g = eng.cluster_lml_grad(slot)
    assert np.max(np.abs(g - c['grad'])) / np.max(np.abs(c['grad'])) < 1e-8
    mu, var = eng.cluster_predict(slot, c['xs'])
    assert np.allclose(mu, c['mu'], rtol=1e-8, atol=1e-10)
    assert np.allclose(var, c['var'], rtol=1e-9)
    lpd = eng.cluster_lpd(slot, c['xs'], c['ys'])
    assert np.max(np.abs(lpd - c['lpd'])) / np.max(np.abs(c['lpd'])) < 1e-9
    f, gt = eng.cluster_objective(slot, c['t'])
    assert f == pytest.approx(float(c['f']), rel=1e-9)
    assert np.allclose(gt, c['g'], rtol=1e-7, atol=1e-9)
    # new-cluster marginal of the query points taken as one patient
    order = np.argsort(c['xs'])
    th0 = np.array([0.0, 1.0, 4.0 if c['spec'][0] == 0 else 1.0, 0.5])
    spec0 = _capi.ModelSpec(int(c['spec'][0]), int(c['spec'][1]), 1, 0)
    ll = eng.new_cluster_ll(spec0, th0, np.array([0, 9]), c['xs'][order], c['ys'], np.array([float(c['a_draw'])]))
    assert ll[0] == pytest.approx(float(c['new_ll']), rel=1e-9)
    eng.close()
