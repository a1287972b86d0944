"""GPU parity of the Gibbs chain: teacher-forced steps (L2), the reference tutorial's golden trace
free-running on the engine (L3), the rank-update fast path against the refactorising oracle, and the
reference's own smoke tests (mogp/test_package.py)."""
import numpy as np
import pytest

import oracle
from conftest import tutorial_data, TUTORIAL_Z0_BLOCK
from test_oracle_kat import TRACE_1_9, Z_FINAL_10

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from mogp_b200 import _capi
    e = _capi.Engine(0)
    yield e
    e.close()


def _theta_of(model):
    th = np.zeros(4)
    th[(0 if model.mean_func else 1):] = model.param_array
    return th


def test_teacher_forced_steps_tutorial(eng, single_blas_thread):
    """Engine fed the oracle's state: same p vector (<=1e-9), same active ids, same counts, every step."""
    from mogp_b200 import MoGP_constrained
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    orc = oracle.MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=3, rand_seed=0, normalize=True, init_z=z0)
    orc.initialize_sampler(orc.num_init_clusters, True)
    orc.allocmodel.calc_suffstats(orc.z)
    orc.step_log = []
    mix = MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=3, rand_seed=0, normalize=True, init_z=z0, engine=eng)
    mix.initialize_sampler(mix.num_init_clusters, True)       # consumes the same 5 randn draws
    worst = 0.0
    for sweep in range(2):
        rs_state = np.random.get_state()
        perm = np.random.permutation(np.arange(60))
        for n in perm:
            # draws the oracle is about to consume, replayed for the engine
            st = np.random.get_state()
            a = np.random.randn(1, 1)[0, 0]
            u = np.random.random_sample()
            np.random.set_state(st)
            orc.gibbs_step(n)
            rec = orc.step_log[-1]
            eng.chain_step_begin(n)
            ids, score = eng.chain_step_score(a)
            from mogp_b200._capi import host_choice
            idx, p, _margin = host_choice(score, u)
            assert list(ids) == list(rec['active'])
            worst = max(worst, np.max(np.abs(p - rec['p'])))
            assert np.max(np.abs(p - rec['p'])) < 1e-9
            assert ids[idx] == rec['z']
            is_new = eng.chain_step_commit(rec['z'], a)
            # teacher forcing: adopt the oracle's post-refit hyper-parameters
            _z, Nk, slots = eng.chain_state(60)
            assert list(Nk) == list(orc.allocmodel.Nk)
            if not is_new:
                eng.cluster_set_theta(int(slots[rec['z']]), _theta_of(orc.obsmodel[rec['z']].model))
                assert abs(eng.cluster_loglik(int(slots[rec['z']])) - orc.obsmodel[rec['z']].model.log_likelihood()) \
                    <= 1e-9 * abs(orc.obsmodel[rec['z']].model.log_likelihood())
    print("worst |dp| over teacher-forced steps:", worst)


def test_tutorial_chain_reproduces_the_notebook(eng):
    """L3: the free-running engine chain prints the reference notebook's trace
    (example/tutorial_train_mogp_model.ipynb:L117-127, L201-203, L272-289)."""
    from mogp_b200 import MoGP_constrained
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    mix = MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=10, rand_seed=0, normalize=True, init_z=z0, engine=eng)
    mix.sample()
    assert list(mix.init_occupancy) == [30, 6, 12, 6, 6]
    assert [list(t) for t in mix.trace] == TRACE_1_9
    assert list(mix.z) == Z_FINAL_10
    assert list(np.where(mix.allocmodel.Nk > 0)[0]) == [0, 1, 3]
    m = mix.obsmodel[0].model
    assert m.param_names == ['neg_linmap.A', 'rbf.variance', 'rbf.lengthscale', 'Gaussian_noise.variance']
    assert m[0] == pytest.approx(0.28341388, abs=5e-8)
    assert m[1] == 1.0
    assert m[2] == pytest.approx(1.5770597538950126, rel=1e-6)
    assert m[3] == pytest.approx(0.013348237935366946, rel=1e-6)
    assert m.objective_function() == pytest.approx(-37.289086611673795, rel=1e-9)
    print("min draw margin:", mix.min_margin)


@pytest.mark.parametrize("kernel,mean_func,threshold", [('rbf', True, 0.5), ('rbf', False, None), ('linear', True, None)])
def test_fast_sweeps_match_oracle(eng, kernel, mean_func, threshold, single_blas_thread):
    """refit='sweep': the engine's rank-n up/down-dates + leave-block-out score against the oracle, which
    refactorises on every move as the reference does. Same assignments, same counts, LMLs to 1e-9."""
    from mogp_b200 import MoGP_constrained
    from mogp_b200.synth import alsfrs_like, block_init
    N = 120
    X, Y = alsfrs_like(N, seed=3)
    z0 = block_init(N, 4, X)
    kw = dict(alpha=0.9, num_iter=4, rand_seed=1, normalize=True, kernel=kernel, mean_func=mean_func,
              threshold=threshold, noise_variance=0.5, refit='sweep', init_z=z0)
    orc = oracle.MoGP_constrained(X=X.copy(), Y=Y.copy(), **kw)
    orc.sample()
    mix = MoGP_constrained(X=X.copy(), Y=Y.copy(), engine=eng, **kw)
    mix.sample()
    print("min draw margin:", mix.min_margin, "traces:", [list(t) for t in mix.trace])
    assert [list(t) for t in mix.trace] == [list(t) for t in orc.trace]
    assert list(mix.z) == list(orc.z)
    assert list(mix.allocmodel.Nk) == list(orc.allocmodel.Nk)
    assert np.allclose(mix.lobs, orc.lobs, rtol=1e-6)
    assert np.allclose(mix.lalloc, orc.lalloc, rtol=1e-13)


@pytest.mark.parametrize("lookahead,pipeline,env", [(16, 2, {}), (32, 2, {}), (48, 2, {}), (64, 2, {}), (32, 1, {}),
                                                    (16, 3, {}), (32, 3, {}), (64, 3, {}),   # two passes in flight
                                                    (48, 3, {"MOGP_TMAINT": "0"}),
                                                    (48, 2, {"MOGP_RANK_RESCORE": "0"}),   # every change re-scored
                                                    (48, 2, {"MOGP_TMAINT": "0"}),         # cached columns not kept current: one correction per cluster and batch, the rest re-scored
                                                    (32, 2, {"MOGP_TMAINT": "0"}),
                                                    (48, 2, {"MOGP_ZIGZAG": "1"}),
                                                    (48, 2, {"MOGP_PRUNE": "0"}),          # threshold rule after the product only
                                                    (64, 2, {"MOGP_CHECK_CORRECTIONS": "1"}),
                                                    (64, 2, {"MOGP_CHECK_CORRECTIONS": "2"}),
                                                    (64, 3, {"MOGP_CHECK_CORRECTIONS": "2"})])
def test_lookahead_window_equals_patient_by_patient(eng, lookahead, pipeline, env, monkeypatch):
    """mogp_chain_sweep with the look-ahead pipeline (batches of patients scored per pass over the factors, the next
    batch's pass running while the current batch moves, cached scores corrected by rank-n formulas or re-scored against
    the clusters that change) walks the same chain as one pass per patient: identical choices per step,
    identical counts, cluster log marginals within 1e-9 relative (after an L-BFGS-B refit in between), on a first sweep where most patients move (every
    window sees invalidations, new clusters and emptied clusters)."""
    from mogp_b200 import MoGP_constrained
    from mogp_b200.synth import alsfrs_like, block_init
    monkeypatch.setenv("MOGP_PIPELINE", str(pipeline))   # passes in flight: 2 = the next batch's overlaps the current batch's moves, 3 = the next two
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    N = 400
    X, Y = alsfrs_like(N, seed=5)
    z0 = block_init(N, 6, X)
    z0[:3] = np.array([6, 7, 7])                     # a singleton and a pair: clusters that empty during the sweep
    runs = []
    for la in (-1, lookahead):
        kw = dict(alpha=2.0, num_iter=3, rand_seed=2, normalize=True, threshold=0.5, noise_variance=0.5,
                  refit='sweep', init_z=z0, lookahead=la)
        mix = MoGP_constrained(X=X.copy(), Y=Y.copy(), engine=eng, **kw)
        mix.initialize_sampler(mix.num_init_clusters, True)
        chosen = []
        for _ in range(2):
            perm, a, u = mix._draws_for_sweep(N)
            c, mm = eng.chain_sweep(perm, a, u)
            chosen.append(np.array(c))
            mix._sync_state()
            lml = {k: mix.obsmodel[k].calc_ll() for k in mix._active()}
            mix.refit_all()
        runs.append((chosen, mix.allocmodel.Nk.copy(), lml, mm))
    (c0, Nk0, l0, mm0), (c1, Nk1, l1, mm1) = runs
    print("min margins:", mm0, mm1, "clusters:", len(l0))
    for a_, b_ in zip(c0, c1):
        assert np.array_equal(a_, b_)
    assert list(Nk0) == list(Nk1)
    assert l0.keys() == l1.keys()
    for k in l0:
        assert abs(l0[k] - l1[k]) <= 1e-9 * max(1.0, abs(l0[k]))


@pytest.mark.parametrize("mode", ["1", "2"])
def test_rank_n_score_corrections_equal_rescoring(eng, monkeypatch, mode):
    """After a move the cached scores of the patients ahead under the two clusters it changed are corrected by rank-n
    formulas (k_alg_leave / k_alg_join) while the factors are still being updated, and the cached columns T = M Kx, ss, mu
    are brought to the changed cluster on the device (k_T_leave, the write-backs of the corrections), for this batch and -
    chained behind its pass - for the next.  In check mode every corrected pair is re-scored from the UPDATED factor:
    log-likelihoods and threshold means agree within 1e-9 relative.  Mode 1: the re-scored columns take over (every
    correction starts from fresh columns); mode 2: compare only - corrections keep building on maintained columns, and the
    scores a batch takes from the corrections chained behind its pass are checked when it starts."""
    from mogp_b200 import MoGP_constrained
    from mogp_b200.synth import alsfrs_like, block_init
    monkeypatch.setenv("MOGP_CHECK_CORRECTIONS", mode)
    N = 600
    X, Y = alsfrs_like(N, seed=11)
    z0 = block_init(N, 5, X)
    mix = MoGP_constrained(X=X.copy(), Y=Y.copy(), engine=eng, alpha=1.5, num_iter=3, rand_seed=4, normalize=True,
                           threshold=0.5, noise_variance=0.5, refit='sweep', init_z=z0)
    mix.initialize_sampler(mix.num_init_clusters, True)
    c0 = eng.chain_counters()
    for _ in range(2):
        mix.sweep()
    c1 = eng.chain_counters()
    checked = c1["corrections_checked"] - c0["corrections_checked"]
    print("pairs checked:", checked, "worst relative disagreement:", c1["max_correction_err"])
    print("batches served by chained corrections:", c1["replayed"] - c0["replayed"], "re-scoring passes:",
          c1["rescoring_passes"] - c0["rescoring_passes"])
    assert c1["rank_corrections"] - c0["rank_corrections"] > 50 and checked > 200
    assert c1["max_correction_err"] <= 1e-9
    if mode == "2":
        assert c1["replayed"] - c0["replayed"] > 20


@pytest.mark.slow
def test_cfg1_reference_schedule_200_patients(eng):
    """BASELINE configs[0] shape (MoGP_constrained, 200 sparse ALSFRS-R-like patients, rbf, threshold 0.5, refit after
    every move = the reference's schedule), free-running for 3 iterations: same occupancy trace, assignments and
    counts as the CPU oracle; hyper-parameters and log-likelihoods agree to the optimiser's tolerance."""
    from mogp_b200 import MoGP_constrained
    from mogp_b200.synth import alsfrs_like, block_init
    N = 200
    X, Y = alsfrs_like(N, seed=0)
    z0 = block_init(N, round(N / 50), X)
    kw = dict(alpha=round(N / 50) / np.log10(N), num_iter=3, rand_seed=0, normalize=True, threshold=0.5,
              noise_variance=0.5, init_z=z0)
    orc = oracle.MoGP_constrained(X=X.copy(), Y=Y.copy(), **kw)
    orc.sample()
    mix = MoGP_constrained(X=X.copy(), Y=Y.copy(), engine=eng, **kw)
    mix.sample()
    print("cfg1: min draw margin", mix.min_margin, "trace", [list(t) for t in mix.trace])
    assert [list(t) for t in mix.trace] == [list(t) for t in orc.trace]
    assert list(mix.z) == list(orc.z)
    assert list(mix.allocmodel.Nk) == list(orc.allocmodel.Nk)
    assert np.allclose(mix.lobs, orc.lobs, rtol=1e-6)
    for k in np.where(orc.allocmodel.Nk > 0)[0]:
        assert np.allclose(mix.obsmodel[k].model.param_array, orc.obsmodel[k].model.param_array, rtol=1e-4)


def test_reference_smoke_nan_rows(eng):
    """mogp/test_package.py:30-46 — NaN in the middle / at the end of a row, one initial cluster."""
    from mogp_b200 import MoGP_constrained
    X = np.array([[0., 1., 2., 3.], [0., 1., np.nan, 3.], [0., 1., 2., np.nan]])
    Y = np.array([[48., 40., 35., 30.], [48., 41., np.nan, 28.], [48., 39., 33., np.nan]])
    mix = MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=5, rand_seed=0, normalize=True, num_init_clusters=1,
                           engine=eng)
    mix.sample()
    assert mix.allocmodel.Nk.sum() == 3
    orc = oracle.MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=5, rand_seed=0, normalize=True, num_init_clusters=1)
    orc.sample()
    assert list(mix.z) == list(orc.z)


def test_predict_and_rank(eng):
    """MoGP_constrained.predict / utils.rank_cluster_prediction (mogp_constrained.py:312-347, utils.py:9-29;
    query of mogp/test_package.py:57-58) against the oracle on the same fitted state."""
    from mogp_b200 import MoGP_constrained, utils
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    kw = dict(alpha=1., num_iter=4, rand_seed=0, normalize=True, init_z=z0)
    orc = oracle.MoGP_constrained(X=X, Y=Y, **kw)
    orc.sample()
    mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
    mix.sample()
    assert list(mix.z) == list(orc.z)
    # same hyper-parameters on both sides, then compare predict()
    _z, _Nk, slots = eng.chain_state(60)
    for k in np.where(orc.allocmodel.Nk > 0)[0]:
        eng.cluster_set_theta(int(slots[k]), _theta_of(orc.obsmodel[k].model))
    x_new, y_new = np.array([0.25, 0.5, 1, 2.4]), np.array([46., 44., 41., 19.])
    np.random.seed(11)
    po = orc.predict(x_new, y_new)
    np.random.seed(11)
    pe = mix.predict(x_new, y_new)
    assert pe.shape == po.shape
    assert np.max(np.abs(pe - po)) < 1e-9
    np.random.seed(11)
    ro = oracle.rank_cluster_prediction(orc, x_new, y_new)
    np.random.seed(11)
    re = utils.rank_cluster_prediction(mix, x_new, y_new)
    assert list(ro[0]) == list(re[0])


def test_pickle_roundtrip(eng, tmp_path):
    """joblib checkpoints (mogp_constrained.py:174-179, 289-310): dump, load, predict from the loaded model."""
    import joblib
    from mogp_b200 import MoGP_constrained
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    mix = MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=3, rand_seed=0, normalize=True, init_z=z0,
                           savepath=tmp_path, savename='m', engine=eng)
    mix.sample()
    assert (tmp_path / 'm_iterinit.pkl').exists() and (tmp_path / 'm.pkl').exists()
    loaded = joblib.load(tmp_path / 'm.pkl')
    k = int(np.where(loaded.allocmodel.Nk > 0)[0][0])
    xs = np.linspace(0, 6, 9).reshape(-1, 1)
    m0, v0 = mix.obsmodel[k].model.predict(xs)
    m1, v1 = loaded.obsmodel[k].model.predict(xs)
    assert np.allclose(m0, m1, rtol=1e-12) and np.allclose(v0, v1, rtol=1e-12)
    assert loaded.obsmodel[k].model.normalizer is not None
