"""GPU parity tests: the CUDA engine (through the C ABI, include/mogp_b200.h) against the CPU oracle on the
same seeded inputs.  Tolerance: 1e-9 relative for log-likelihoods (BASELINE.json north_star), bit-exact for
counts / assignments."""
import numpy as np
import pytest

import oracle
from oracle.gpy_restated import GPRegressionOracle

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from mogp_b200 import _capi
    e = _capi.Engine(0)
    yield e
    e.close()


def make_cluster(rs, m, lo=0.0, hi=8.0):
    x = np.sort(rs.uniform(lo, hi, m))
    x[: max(1, m // 8)] = 0.0                      # onset anchors: repeated inputs
    y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
    return x, y


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


CASES = [  # kernel, mean_func, sig_fix, noise_fix, theta (A, k0, k1, noise), m
    ('rbf', True, True, False, (0.28, 1.0, 1.58, 0.0133), 37),
    ('rbf', True, True, False, (0.5, 1.0, 4.0, 0.5), 64),
    ('rbf', True, False, False, (0.1, 0.7, 2.5, 0.2), 65),
    ('rbf', False, True, False, (0.0, 1.0, 1.0, 0.05), 200),
    ('rbf', True, True, True, (0.3, 1.0, 3.0, 0.3), 450),
    ('linear', True, True, False, (0.4, 1.0, 1.0, 0.5), 130),
    ('linear', False, False, False, (0.0, 0.6, 0.8, 0.1), 70),
    ('rbf', True, True, False, (0.3, 1.0, 2.0, 0.02), 1100),
]


def build_pair(eng, case, seed=0):
    from mogp_b200.obsmodel import make_spec
    kernel, mf, sfix, nfix, th, m = case
    rs = np.random.RandomState(seed + m)
    x, y = make_cluster(rs, m)
    o = GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name=kernel, signal_variance=th[1],
                           signal_variance_fix=sfix, noise_variance=th[3], noise_variance_fix=nfix, mean_func=mf,
                           lengthscale=th[2], a_draw=th[0] if mf else None)
    if kernel == 'linear':
        o.param_array[o._off + 1] = th[2]
        o._inference()
    spec = make_spec(kernel, mf, sfix, nfix)
    slot = eng.cluster_create(spec, np.array(th))
    eng.cluster_set_data(slot, x, y)
    return o, slot, spec, rs


@pytest.mark.parametrize("case", CASES, ids=lambda c: "{}-mf{}-m{}".format(c[0], int(c[1]), c[5]))
def test_inference_lml_grad_objective(eng, case):
    o, slot, spec, rs = build_pair(eng, case)
    assert rel(eng.cluster_loglik(slot), o.log_likelihood()) < RTOL
    g = eng.cluster_lml_grad(slot)
    go = np.zeros(4)
    go[(0 if case[1] else 1):] = o._grad_lml
    scale = np.max(np.abs(go))
    assert np.max(np.abs(g - go)) / scale < 1e-8
    # optimiser objective (prior + Jacobian included) at a perturbed point
    t = o.optimizer_array() + rs.uniform(-0.3, 0.3, o.optimizer_array().size)
    fo, dfo = o.objective_grads(t)
    f, df = eng.cluster_objective(slot, t)
    assert rel(f, fo) < RTOL
    assert np.max(np.abs(df - dfo)) / max(np.max(np.abs(dfo)), 1e-12) < 1e-8
    assert np.allclose(eng.cluster_optimizer_array(slot), t, rtol=1e-12, atol=1e-12)
    eng.cluster_free(slot)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "{}-mf{}-m{}".format(c[0], int(c[1]), c[5]))
def test_predict_and_density(eng, case):
    o, slot, spec, rs = build_pair(eng, case, seed=1)
    for n in (1, 4, 11, 29, 80):
        xs = rs.uniform(0, 9, n)
        ys = rs.randn(n)
        mu, var = eng.cluster_predict(slot, xs)
        mo, vo = o.predict(xs.reshape(-1, 1))
        assert rel(mu, mo.ravel()) < 1e-8 or np.max(np.abs(mu - mo.ravel())) < 1e-10
        assert rel(var, vo.ravel()) < RTOL
        lpd = eng.cluster_lpd(slot, xs, ys)
        lo = o.log_predictive_density(xs.reshape(-1, 1), ys.reshape(-1, 1)).ravel()
        assert np.max(np.abs(lpd - lo)) / np.max(np.abs(lo)) < RTOL
    eng.cluster_free(slot)


def test_score_pairs_many_clusters(eng):
    from mogp_b200.synth import alsfrs_like
    from mogp_b200.utils import pack_patients
    pairs = [build_pair(eng, c, seed=2) for c in CASES[:5]]
    slots = [p[1] for p in pairs]
    X, Y = alsfrs_like(23, seed=5)
    Y = (Y - 30) / 20
    X[3, 4:] = np.nan
    Y[3, 4:] = np.nan                                          # a ragged short row
    off, x, y = pack_patients(X, Y)
    score, mu = eng.score_pairs(off, x, y, slots, thr_index=1, want_mu=True)
    for p in range(X.shape[0]):
        xn, yn = x[off[p]:off[p + 1]].reshape(-1, 1), y[off[p]:off[p + 1]].reshape(-1, 1)
        for c, (o, *_r) in enumerate(pairs):
            so = np.sum(o.log_predictive_density(xn, yn))
            assert abs(score[p, c] - so) / abs(so) < RTOL
            assert abs(mu[p, c] - o.predict(xn[1].reshape(-1, 1))[0][0, 0]) < 1e-9
    for s in slots:
        eng.cluster_free(s)


@pytest.mark.parametrize("kernel,mf", [('rbf', True), ('rbf', False), ('linear', True)])
def test_new_cluster_ll(eng, kernel, mf):
    from mogp_b200.obsmodel import make_spec, initial_theta
    from mogp_b200.synth import alsfrs_like
    from mogp_b200.utils import pack_patients
    X, Y = alsfrs_like(40, seed=9)
    Y = (Y - 30) / 20
    off, x, y = pack_patients(X, Y)
    rs = np.random.RandomState(3)
    a = rs.randn(40)
    spec = make_spec(kernel, mf, True, False)
    th = initial_theta(kernel, 1.0, 0.5, 4.0, 0.0)
    ll = eng.new_cluster_ll(spec, th, off, x, y, a if mf else None)
    for p in range(40):
        xn, yn = x[off[p]:off[p + 1]].reshape(-1, 1), y[off[p]:off[p + 1]].reshape(-1, 1)
        o = GPRegressionOracle(xn, yn, kernel_name=kernel, signal_variance=1.0, noise_variance=0.5, mean_func=mf,
                               a_draw=a[p] if mf else None)
        assert abs(ll[p] - o.log_likelihood()) / abs(o.log_likelihood()) < RTOL


@pytest.mark.parametrize("m,case", [(150, CASES[0]), (700, CASES[1]), (333, CASES[5])])
def test_rank_updates_match_refactorisation(eng, m, case):
    """append / remove_rows / leave_out_score at fixed theta == the reference's full set_XY on the new data."""
    from mogp_b200.obsmodel import make_spec
    kernel, mf, sfix, nfix, th, _ = case
    rs = np.random.RandomState(m)
    x, y = make_cluster(rs, m)
    spec = make_spec(kernel, mf, sfix, nfix)
    slot = eng.cluster_create(spec, np.array(th))
    eng.cluster_set_data(slot, x, y)

    def oracle_for(xx, yy):
        o = GPRegressionOracle(xx.reshape(-1, 1), yy.reshape(-1, 1), kernel_name=kernel, signal_variance=th[1],
                               signal_variance_fix=sfix, noise_variance=th[3], noise_variance_fix=nfix,
                               mean_func=mf, lengthscale=th[2], a_draw=th[0] if mf else None)
        if kernel == 'linear':
            o.param_array[o._off + 1] = th[2]
            o._inference()
        return o

    xq, yq = rs.uniform(0, 8, 7), rs.randn(7)
    cur_x, cur_y = x.copy(), y.copy()
    for it, (p, n) in enumerate([(5, 7), (60, 11), (m - 30, 4), (0, 3), (64, 16), (100, 1)]):
        # leave-out score of rows [p, p+n) vs oracle on the reduced data
        s, mu = eng.cluster_leave_out_score(slot, p, n, thr_index=min(1, n - 1))
        keep = np.r_[0:p, p + n:cur_x.size]
        o_red = oracle_for(cur_x[keep], cur_y[keep])
        xb, yb = cur_x[p:p + n].reshape(-1, 1), cur_y[p:p + n].reshape(-1, 1)
        so = np.sum(o_red.log_predictive_density(xb, yb))
        assert abs(s - so) / abs(so) < RTOL, "leave-out score, it {}".format(it)
        assert abs(mu - o_red.predict(xb[min(1, n - 1)].reshape(-1, 1))[0][0, 0]) < 1e-9
        # actual removal
        eng.cluster_remove_rows(slot, p, n)
        assert eng.cluster_size(slot) == keep.size
        assert rel(eng.cluster_loglik(slot), o_red.log_likelihood()) < RTOL
        sc = eng.score_pairs(np.array([0, 7]), xq, yq, [slot])[0, 0]
        assert abs(sc - np.sum(o_red.log_predictive_density(xq.reshape(-1, 1), yq.reshape(-1, 1)))) / abs(sc) < RTOL
        # append them back at the end
        eng.cluster_append(slot, xb.ravel(), yb.ravel())
        cur_x, cur_y = np.r_[cur_x[keep], xb.ravel()], np.r_[cur_y[keep], yb.ravel()]
        o_full = oracle_for(cur_x, cur_y)
        assert rel(eng.cluster_loglik(slot), o_full.log_likelihood()) < RTOL
        sc = eng.score_pairs(np.array([0, 7]), xq, yq, [slot])[0, 0]
        assert abs(sc - np.sum(o_full.log_predictive_density(xq.reshape(-1, 1), yq.reshape(-1, 1)))) / abs(sc) < RTOL
        gx, gy = eng.cluster_get_data(slot)
        assert np.array_equal(gx, cur_x) and np.array_equal(gy, cur_y)
    # a long append crossing several tiles
    xa, ya = make_cluster(rs, 150)
    eng.cluster_append(slot, xa, ya)
    o_full = oracle_for(np.r_[cur_x, xa], np.r_[cur_y, ya])
    assert rel(eng.cluster_loglik(slot), o_full.log_likelihood()) < RTOL
    g = eng.cluster_lml_grad(slot)
    go = np.zeros(4)
    go[(0 if mf else 1):] = o_full._grad_lml
    assert np.max(np.abs(g - go)) / np.max(np.abs(go)) < 1e-8
    eng.cluster_free(slot)


def test_refit_reaches_the_oracle_optimum(eng):
    case = CASES[0]
    o, slot, spec, rs = build_pair(eng, case, seed=4)
    from mogp_b200.obsmodel import GPModel
    gm = GPModel(eng, spec, None, slot=slot)
    o.optimize()
    gm.optimize()
    th = eng.cluster_get_theta(slot)
    assert np.allclose(th, o.param_array, rtol=2e-6)
    assert rel(gm.objective_function(), o.objective()) < 1e-9
    eng.cluster_free(slot)


@pytest.mark.parametrize("case", [('rbf', True, True, False, (0.3, 1.0, 3.0, 0.3), 300),
                                  ('rbf', False, False, False, (0.0, 0.8, 2.0, 0.1), 520),
                                  ('linear', True, True, False, (0.4, 1.0, 1.0, 0.5), 260)],
                         ids=lambda c: "{}-mf{}-m{}".format(c[0], int(c[1]), c[5]))
def test_refit_with_repeated_inputs_collapsed(eng, case, monkeypatch):
    """Clusters of >= 128 points with > 5 % repeated inputs (every patient's onset anchor): the native refit evaluates
    the optimiser's points on a collapsed shadow (refit.inl: one weighted observation per distinct input + a closed-form
    factor) and the closing f(x_opt) on the cluster itself.  Same optimum, same objective, same number of evaluations
    as the oracle's SciPy L-BFGS-B on the full data; the cluster is left with a valid factor at the optimum.
    (MOGP_WS_REFIT=0: RBF clusters would otherwise be evaluated in weight space, test_refit_in_weight_space.)"""
    monkeypatch.setenv("MOGP_WS_REFIT", "0")
    o, slot, spec, rs = build_pair(eng, case, seed=11)
    n0 = o.n_evals
    o.optimize()
    c0 = eng.refit_counters()
    evals = eng.refit_clusters([slot])
    c1 = eng.refit_counters()
    assert c1["launched"] > c0["launched"]
    th = eng.cluster_get_theta(slot)
    assert np.allclose(th[(0 if case[1] else 1):], o.param_array, rtol=2e-6)   # (no slope parameter without the mean function)
    assert rel(eng.cluster_loglik(slot), o.log_likelihood()) < 1e-9
    assert int(evals[0]) == o.n_evals - n0          # the optimiser's evaluations + the closing f(x_opt)
    xs = rs.uniform(0, 8, 5)
    mu, var = eng.cluster_predict(slot, xs)
    mo, vo = o.predict(xs.reshape(-1, 1))
    assert rel(mu, mo.ravel()) < 1e-7 and rel(var, vo.ravel()) < 1e-7
    eng.cluster_free(slot)


@pytest.mark.parametrize("case", [('rbf', True, True, False, (0.3, 1.0, 3.0, 0.3), 300),
                                  ('rbf', False, False, False, (0.0, 0.8, 2.0, 0.1), 520),
                                  ('rbf', True, True, False, (0.28, 1.0, 1.58, 0.0133), 1100)],
                         ids=lambda c: "{}-mf{}-m{}".format(c[0], int(c[1]), c[5]))
def test_refit_in_weight_space(eng, case, monkeypatch):
    """MOGP_WS_REFIT=1: the optimiser's evaluations of an RBF cluster run in weight space (wspace.cuh: the kernel
    interpolated at 48 Chebyshev nodes, O(r^3) per evaluation whatever m; the closing f(x_opt) on the cluster itself).  Same
    optimum, same objective, same number of evaluations as the oracle's SciPy L-BFGS-B on the dense GP, and the cluster is
    left with a valid dense factor at the optimum."""
    monkeypatch.setenv("MOGP_WS_REFIT", "1")
    o, slot, spec, rs = build_pair(eng, case, seed=13)
    n0 = o.n_evals
    o.optimize()
    c0 = eng.refit_counters()
    evals = eng.refit_clusters([slot])
    c1 = eng.refit_counters()
    assert c1["launched"] > c0["launched"]
    th = eng.cluster_get_theta(slot)
    assert np.allclose(th[(0 if case[1] else 1):], o.param_array, rtol=2e-6)
    assert rel(eng.cluster_loglik(slot), o.log_likelihood()) < 1e-9
    assert abs(int(evals[0]) - (o.n_evals - n0)) <= 1
    xs = rs.uniform(0, 8, 5)
    mu, var = eng.cluster_predict(slot, xs)
    mo, vo = o.predict(xs.reshape(-1, 1))
    assert rel(mu, mo.ravel()) < 1e-7 and rel(var, vo.ravel()) < 1e-7
    eng.cluster_free(slot)


def test_worker_engine_objective_and_concurrent_refits(eng):
    """mogp_cluster_objective_on: a cluster of one engine evaluated on another engine's stream - first at the
    cluster's CURRENT theta (the owner's factor is reused, only the gradient runs on the worker), then at a new one;
    and `optimize_many` (concurrent refits on worker engines) lands where sequential `optimize()` does."""
    from mogp_b200 import _capi
    from mogp_b200.obsmodel import GPModel, optimize_many
    worker = _capi.Engine(0, blocking_sync=True)
    pairs = [build_pair(eng, c, seed=7) for c in (CASES[0], CASES[2], CASES[5])]
    for o, slot, spec, rs in pairs:
        t0 = eng.cluster_optimizer_array(slot)
        f_w, g_w = eng.cluster_objective(slot, t0, worker=worker)          # same theta: no re-inference
        f_o, g_o = o.objective_grads(o.optimizer_array())
        assert rel(f_w, f_o) < RTOL and np.max(np.abs(g_w - g_o)) / np.max(np.abs(g_o)) < 1e-8
        t1 = t0 + 0.05
        f_w, g_w = eng.cluster_objective(slot, t1, worker=worker)
        f_o, g_o = o.objective_grads(t1)
        assert rel(f_w, f_o) < RTOL and np.max(np.abs(g_w - g_o)) / np.max(np.abs(g_o)) < 1e-8
    models = [GPModel(eng, spec, None, slot=slot) for _o, slot, spec, _r in pairs]
    start = [eng.cluster_get_theta(m.slot) for m in models]
    optimize_many(models, [worker, _capi.Engine(0)])
    conc = [eng.cluster_get_theta(m.slot) for m in models]
    for m, th in zip(models, start):
        eng.cluster_set_theta(m.slot, th)
        m.optimize()
    for m, th in zip(models, conc):
        assert np.array_equal(eng.cluster_get_theta(m.slot), th)           # bit-identical: same evaluations
    for _o, slot, _s, _r in pairs:
        eng.cluster_free(slot)
    worker.close()
