"""GPU edge cases: ragged / single-visit / long rows, emptied and new clusters, alpha extremes, capacity growth,
and size-independent round-trip properties at the headline cluster size (m ~ 2 500)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from mogp_b200 import _capi
    e = _capi.Engine(0)
    yield e
    e.close()


def ragged_data(seed, N=40, T=24):
    rs = np.random.RandomState(seed)
    X = np.full((N, T), np.nan)
    Y = np.full((N, T), np.nan)
    for i in range(N):
        n = [1, 2, 23][i % 3] if i < 9 else rs.randint(3, 9)      # single-visit, two-visit and 23-visit rows
        t = np.sort(rs.uniform(0.2, 8, n))
        X[i, 0], Y[i, 0] = 0.0, 48.0
        X[i, 1:1 + n] = t
        Y[i, 1:1 + n] = np.clip(np.round(48 - rs.uniform(2, 9) * t + rs.randn(n) * 2), 0, 48)
    holes = rs.randint(9, N, 6)                                    # NaN in the middle of a row
    for i in holes:
        X[i, 2], Y[i, 2] = np.nan, np.nan
    return X, Y


@pytest.mark.parametrize("alpha", [0.0, 1.0, 50.0])
def test_ragged_rows_and_alpha_extremes(eng, alpha, single_blas_thread):
    """Rows with 1, 2 and 23 visits (> the 16-row rank-update limit: those patients take the refactorising path),
    NaN holes, alpha = 0 (new clusters all but impossible) and alpha = 50 (many new and emptied clusters)."""
    from mogp_b200 import MoGP_constrained
    X, Y = ragged_data(1)
    z0 = (np.arange(40) % 3).astype(np.int32)
    kw = dict(alpha=alpha, num_iter=3, rand_seed=2, normalize=True, threshold=0.5, noise_variance=0.5, refit='sweep',
              init_z=z0)
    orc = oracle.MoGP_constrained(X=X.copy(), Y=Y.copy(), **kw)
    orc.sample()
    mix = MoGP_constrained(X=X.copy(), Y=Y.copy(), engine=eng, **kw)
    mix.sample()
    assert list(mix.z) == list(orc.z)
    assert list(mix.allocmodel.Nk) == list(orc.allocmodel.Nk)
    assert mix.allocmodel.Nk.dtype == np.int64
    if alpha == 50.0:
        assert len(mix.allocmodel.Nk) > 3 and (mix.allocmodel.Nk == 0).any()      # ids are never reused
        assert all(mix.obsmodel[k] is None for k in np.where(mix.allocmodel.Nk == 0)[0])
    assert np.allclose(mix.lobs, orc.lobs, rtol=1e-6)


def test_capacity_growth_by_appends(eng):
    """A cluster grown patient by patient across several capacity classes equals one built in one go."""
    from mogp_b200.obsmodel import make_spec
    rs = np.random.RandomState(3)
    spec = make_spec('rbf', True, True, False)
    th = np.array([0.3, 1.0, 2.0, 0.1])
    x = rs.uniform(0, 8, 420)
    y = (48 - 6 * x + rs.randn(420) * 3 - 30) / 20
    a = eng.cluster_create(spec, th)
    eng.cluster_set_data(a, x[:5], y[:5])
    pos = 5
    while pos < 420:
        n = int(rs.randint(1, 12))
        eng.cluster_append(a, x[pos:pos + n], y[pos:pos + n])
        pos += n
    b = eng.cluster_create(spec, th)
    eng.cluster_set_data(b, x, y)
    assert eng.cluster_size(a) == 420
    assert abs(eng.cluster_loglik(a) - eng.cluster_loglik(b)) <= 1e-9 * abs(eng.cluster_loglik(b))
    Ma, aa = eng.cluster_get_factor(a)
    Mb, ab = eng.cluster_get_factor(b)
    assert np.max(np.abs(Ma - Mb)) / np.max(np.abs(Mb)) < 1e-9
    assert np.max(np.abs(aa - ab)) / np.max(np.abs(ab)) < 1e-8
    eng.cluster_free(a)
    eng.cluster_free(b)


def test_round_trips_at_headline_cluster_size(eng):
    """m = 2 500 (BASELINE's 10k-patient configuration): properties that need no oracle run.
    (1) leave-block-out score == score after really deleting the rows; (2) delete + append restores the log marginal;
    (3) the log marginal is invariant to the order of the data; (4) the factor inverts: M K_y M^T = I on a probe."""
    from mogp_b200.obsmodel import make_spec
    rs = np.random.RandomState(11)
    m = 2500
    x = np.sort(rs.uniform(0, 8, m))
    x[:300] = 0.0
    y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
    spec = make_spec('rbf', True, True, False)
    th = np.array([0.3, 1.0, 1.6, 0.05])
    s = eng.cluster_create(spec, th)
    eng.cluster_set_data(s, x, y)
    lml0 = eng.cluster_loglik(s)
    p, n = 1237, 9
    xb, yb = x[p:p + n].copy(), y[p:p + n].copy()
    loo, mu = eng.cluster_leave_out_score(s, p, n, thr_index=1)
    eng.cluster_remove_rows(s, p, n)
    sc, mus = eng.score_pairs(np.array([0, n]), xb, yb, [s], thr_index=1, want_mu=True)
    assert abs(loo - sc[0, 0]) <= 1e-9 * abs(sc[0, 0])
    assert abs(mu - mus[0, 0]) <= 1e-9
    eng.cluster_append(s, xb, yb)
    assert abs(eng.cluster_loglik(s) - lml0) <= 1e-10 * abs(lml0)
    perm = rs.permutation(m)
    s2 = eng.cluster_create(spec, th)
    eng.cluster_set_data(s2, x[perm], y[perm])
    assert abs(eng.cluster_loglik(s2) - lml0) <= 1e-10 * abs(lml0)
    # probe: M (K + (noise + 1e-8) I) M^T e = e for random e, with K rebuilt on the host
    M, _alpha = eng.cluster_get_factor(s2)
    xs = x[perm]
    r2 = np.clip(-2 * np.outer(xs, xs) + (xs * xs)[:, None] + (xs * xs)[None, :], 0, None)
    Ky = th[1] * np.exp(-0.5 * r2 / th[2] ** 2) + np.eye(m) * (th[3] + 1e-8)
    e = rs.randn(m)
    out = M @ (Ky @ (M.T @ e))
    assert np.max(np.abs(out - e)) < 1e-8
    eng.cluster_free(s)
    eng.cluster_free(s2)


def test_predict_batch_equals_predict_loop(eng):
    """Batch prediction (patients x clusters in one pass) == the reference's one-patient predict() in a loop."""
    from mogp_b200 import MoGP_constrained
    from mogp_b200.synth import alsfrs_like
    X, Y, z0 = alsfrs_like(90, seed=4, n_classes=4, return_labels=True)
    z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
    mix = MoGP_constrained(X=X, Y=Y, alpha=1.0, num_iter=2, rand_seed=0, normalize=True, noise_variance=0.5,
                           refit='sweep', init_z=z0, engine=eng)
    mix.sample()
    Xq, Yq = alsfrs_like(17, seed=9)
    a = np.random.RandomState(5).randn(17)
    P = mix.predict_batch(Xq, Yq, a)
    assert P.shape == (17, len(mix.allocmodel.Nk) + 1) and np.allclose(P.sum(1), 1.0, atol=1e-12)
    for i in range(17):
        keep = ~np.isnan(Xq[i])
        np.random.seed(123)
        state = np.random.get_state()
        # predict() draws its own randn for the newcomer: replay it so the two paths see the same A
        one = mix.predict_batch(Xq[i:i + 1], Yq[i:i + 1], a[i:i + 1])[0]
        assert np.max(np.abs(one - P[i])) < 1e-12
        np.random.set_state(state)
        p_single = mix.predict(Xq[i][keep], Yq[i][keep])
        assert p_single.shape == P[i].shape and abs(p_single.sum() - 1.0) < 1e-12


def test_full_size_sweep_of_rank_updates_matches_refactorisation(eng):
    """BASELINE's 10k-patient configuration: after a whole sweep (10 000 steps, ~1 700 row deletions + appends per
    sweep at fixed theta, no refit in between) every cluster's log marginal and a probe score equal those of a fresh
    factorisation of the same data to 1e-10 - the rank-n updates do not drift."""
    from mogp_b200 import MoGP_constrained
    from mogp_b200.synth import alsfrs_like
    N, K0 = 10000, 30
    X, Y, z0 = alsfrs_like(N, seed=0, n_classes=K0, return_labels=True)
    z0 = np.unique(z0, return_inverse=True)[1].astype(np.int32)
    mix = MoGP_constrained(X=X, Y=Y, alpha=K0 / np.log10(N), num_iter=3, rand_seed=0, normalize=True, threshold=0.5,
                           noise_variance=0.5, refit='sweep', init_z=z0, engine=eng)
    mix.initialize_sampler(K0, True)
    mix.refit_all()
    perm, a, u = mix._draws_for_sweep(N)
    eng.chain_sweep(perm, a, u)
    mix._sync_state()
    assert mix.allocmodel.Nk.sum() == N and int(np.sum(mix.z != z0)) > 500
    rs = np.random.RandomState(1)
    xq, yq = np.sort(np.r_[0.0, rs.uniform(0.3, 8, 8)]), rs.randn(9)
    for k in mix._active():
        s = mix.obsmodel[k].model.slot
        lml_u = eng.cluster_loglik(s)
        sc_u = eng.score_pairs(np.array([0, 9]), xq, yq, [s])[0, 0]
        x, y = eng.cluster_get_data(s)
        members_visits = sum(int((~np.isnan(X[i])).sum()) for i in np.where(mix.z == k)[0])
        assert x.size == members_visits                      # the factor holds exactly the members' visits
        f = eng.cluster_create(mix._spec(), eng.cluster_get_theta(s))
        eng.cluster_set_data(f, x, y)
        assert abs(lml_u - eng.cluster_loglik(f)) <= 1e-10 * abs(lml_u)
        sc_f = eng.score_pairs(np.array([0, 9]), xq, yq, [f])[0, 0]
        assert abs(sc_u - sc_f) <= 1e-10 * abs(sc_f)
        eng.cluster_free(f)


def test_patient_table_built_on_the_device_equals_numpy_compaction(eng):
    """mogp_set_patients_matrix compacts the reference's NaN-padded N x T matrices on the device
    (mogp_constrained.py:246-247): the CSR it builds equals utils.pack_patients bit for bit - ragged rows, NaN in the
    middle of a row, empty rows, one column, many rows (the scan walks 1024-row chunks) - and a row whose X and Y
    disagree in their NaN pattern is refused like the host packer refuses it."""
    from mogp_b200.utils import pack_patients
    rs = np.random.RandomState(7)
    for N, T in ((1, 1), (3, 4), (1025, 11), (5000, 11), (257, 23)):
        X = rs.uniform(0, 8, (N, T))
        Y = rs.randn(N, T)
        drop = rs.uniform(size=(N, T)) < 0.35
        if N > 2:
            drop[1, :] = True                                   # a patient without visits
            drop[2, :] = False
        X[drop] = np.nan
        Y[drop] = np.nan
        off, x, y = eng.set_patients_matrix(X, Y, thr_col=1)
        off0, x0, y0 = pack_patients(X, Y)
        assert np.array_equal(off, off0)
        assert x.tobytes() == x0.tobytes() and y.tobytes() == y0.tobytes()
    X = np.array([[0., 1., np.nan], [0., 1., 2.]])
    Y = np.array([[48., np.nan, np.nan], [48., 40., 30.]])
    with pytest.raises(ValueError, match="different numbers of non-NaN cells"):
        eng.set_patients_matrix(X, Y)
    with pytest.raises(ValueError, match="different numbers of non-NaN cells"):
        pack_patients(X, Y)


def test_two_models_in_one_process_do_not_share_an_engine():
    """The reference's tutorial trains a constrained and an unconstrained model in one process.  A model built without an
    explicit engine gets its OWN engine (an engine holds one chain): training B must not change what A predicts."""
    from mogp_b200 import MoGP_constrained, MoGP
    from mogp_b200.synth import alsfrs_like, block_init
    X, Y = alsfrs_like(60, seed=13)
    z0 = block_init(60, 3, X)
    kw = dict(alpha=1.0, num_iter=3, rand_seed=1, normalize=True, noise_variance=0.5, refit='sweep', init_z=z0)
    A = MoGP_constrained(X=X.copy(), Y=Y.copy(), threshold=0.5, **kw)
    A.sample()
    xs = np.linspace(0.0, 6.0, 7).reshape(-1, 1)
    before = {int(k): A.obsmodel[k].model.predict(xs) for k in A._active()}
    x_new, y_new = np.array([0.0, 0.5, 1.0, 2.4]), np.array([48., 44., 41., 19.])
    np.random.seed(5)
    p_before = A.predict(x_new, y_new)
    B = MoGP(X=X.copy(), Y=Y.copy(), **kw)
    B.sample()
    assert A.engine is not B.engine
    for k, (m0, v0) in before.items():
        m1, v1 = A.obsmodel[k].model.predict(xs)
        assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    np.random.seed(5)
    assert np.array_equal(A.predict(x_new, y_new), p_before)


def test_stale_cluster_handles_raise(eng):
    """A slot index is reused after a cluster is freed; the handle carries the slot's generation, so a view of the cluster
    that used to live there raises instead of reading its successor."""
    from mogp_b200 import _capi
    from mogp_b200.obsmodel import make_spec
    spec = make_spec('rbf', True, True, False)
    th = np.array([0.3, 1.0, 2.0, 0.3])
    x = np.linspace(0, 5, 20)
    s1 = eng.cluster_create(spec, th)
    eng.cluster_set_data(s1, x, np.sin(x))
    ll1 = eng.cluster_loglik(s1)
    eng.cluster_free(s1)
    s2 = eng.cluster_create(spec, th)
    eng.cluster_set_data(s2, x, np.cos(x))
    assert (s1 & 0xFFFF) == (s2 & 0xFFFF) and s1 != s2            # same slot index, new generation
    with pytest.raises(_capi.EngineError, match="stale cluster handle"):
        eng.cluster_loglik(s1)
    assert eng.cluster_loglik(s2) != ll1
    eng.cluster_free(s2)


def test_rows_longer_than_the_fast_paths(eng):
    """The rank-update fast path takes rows of up to 16 visits, the one-patient marginal kernel up to 64; longer rows go
    through the general cluster path (the reference accepts any row length) - same assignments as the oracle."""
    import oracle
    from mogp_b200 import MoGP_constrained
    rs = np.random.RandomState(3)
    N, T = 12, 80
    X = np.full((N, T), np.nan)
    Y = np.full((N, T), np.nan)
    for i, nv in enumerate([5, 70, 8, 20, 6, 75, 9, 4, 30, 7, 66, 5]):
        t = np.sort(rs.uniform(0.1, 8.0, nv))
        X[i, 0], Y[i, 0] = 0.0, 48.0
        X[i, 1:nv + 1] = t
        Y[i, 1:nv + 1] = np.clip(np.round(48 - (3 + 2 * (i % 3)) * t + rs.randn(nv) * 1.5), 0, 48)
    z0 = (np.arange(N) % 2).astype(np.int32)
    for refit in ('sweep', 'move'):
        kw = dict(alpha=1.0, num_iter=3, rand_seed=2, normalize=True, noise_variance=0.5, threshold=0.5, refit=refit, init_z=z0)
        orc = oracle.MoGP_constrained(X=X.copy(), Y=Y.copy(), **kw)
        orc.sample()
        mix = MoGP_constrained(X=X.copy(), Y=Y.copy(), engine=eng, **kw)
        mix.sample()
        assert list(mix.z) == list(orc.z) and list(mix.allocmodel.Nk) == list(orc.allocmodel.Nk)
        assert np.allclose(mix.lobs, orc.lobs, rtol=1e-6)
    x_new = np.sort(rs.uniform(0.1, 8, 70))
    y_new = np.clip(48 - 5 * x_new, 0, 48)
    np.random.seed(9)
    po = orc.predict(x_new, y_new)
    np.random.seed(9)
    pe = mix.predict(x_new, y_new)
    assert np.max(np.abs(pe - po)) < 1e-8
