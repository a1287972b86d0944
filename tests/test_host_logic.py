"""Host-side logic against NumPy / the oracle: the draw (np.random.choice semantics), CRP bookkeeping,
NaN compaction, argument validation messages, the RNG ledger."""
import numpy as np
import pytest
from scipy.special import logsumexp

import oracle
from mogp_b200 import _capi
from mogp_b200.allocmodel import DPmix
from mogp_b200.utils import pack_patients, generate_toy_data


def numpy_choice_index(p, u):
    """np.random.choice(a, p=p) with its uniform draw replaced by u (numpy/random/mtrand.pyx: choice)."""
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side='right'))


def test_host_choice_matches_numpy_choice():
    rs = np.random.RandomState(0)
    for trial in range(3000):
        n = rs.randint(2, 130)
        score = rs.randn(n) * rs.choice([0.1, 3.0, 40.0])
        if trial % 5 == 0:
            score[rs.randint(n)] = -np.inf              # a threshold-rejected cluster
        u = rs.random_sample()
        idx, p, margin = _capi.host_choice(score, u)
        p_ref = np.exp(score - logsumexp(score))
        assert np.max(np.abs(p - p_ref)) < 1e-12      # exp() of differences up to ~1e2 in magnitude
        ref_idx = numpy_choice_index(p_ref, u)
        if margin > 1e-12:
            assert idx == ref_idx
        assert p[np.isneginf(score)].sum() == 0.0
    with pytest.raises(ValueError):
        _capi.host_choice(np.array([0.0, np.nan]), 0.5)


def test_dpmix_bit_exact_with_oracle():
    rs = np.random.RandomState(1)
    z = rs.randint(0, 7, 200)
    a, b = DPmix(0.75), oracle.DPmix(0.75)
    a.N = b.N = 200
    a.calc_suffstats(z); b.calc_suffstats(z)
    assert a.Nk.dtype == np.int64 and np.array_equal(a.Nk, b.Nk)
    for k in (3, 0, 6):
        sa, ia = a.calc_score(k); sb, ib = b.calc_score(k)
        assert np.array_equal(sa, sb) and np.array_equal(ia, ib)
        a.update_suffstats(len(a.Nk)); b.update_suffstats(len(b.Nk))
        assert np.array_equal(a.Nk, b.Nk)
    assert a.calc_ll() == b.calc_ll()


def test_pack_patients_is_row_major_nan_compaction():
    X = np.array([[0., 1., np.nan, 3.], [0., np.nan, np.nan, np.nan], [0., 1., 2., 3.]])
    Y = X * 2 + 1
    off, x, y = pack_patients(X, Y)
    assert list(off) == [0, 3, 4, 8]
    assert list(x) == [0., 1., 3., 0., 0., 1., 2., 3.]
    assert np.array_equal(y, x * 2 + 1)
    with pytest.raises(ValueError):
        pack_patients(X, np.where(np.isnan(Y), 0.0, Y))


def test_toy_data_matches_oracle_generator():
    Xa, Ya = generate_toy_data(seed=0)
    Xb, Yb = oracle.generate_toy_data(seed=0)
    assert np.array_equal(Xa, Xb) and np.array_equal(Ya, Yb)


@pytest.mark.parametrize("kw,msg", [
    (dict(alpha=-1.), "parameter alpha must be a positive number"),
    (dict(alpha=1., num_init_clusters=0), "number of initial clusters"),
    (dict(alpha=1., num_iter=0), "number of iterations"),
    (dict(alpha=1., kernel='matern'), "Implemented kernels limited to"),
    (dict(alpha=1., signal_variance='a'), "parameter signal_variance must be number"),
    (dict(alpha=1., mean_func=1), "parameter mean_func must be boolean"),
    (dict(alpha=1., threshold='x'), "parameter threshold must be None or number"),
    (dict(alpha=1., normalize=True, Y_mean=1.), "missing std"),
    (dict(alpha=1., Y_mean=1., Y_std=2.), "mean and std cannot be provided"),
])
def test_validation_messages_match_reference(kw, msg):
    """mogp_constrained.py:92-144 — same ValueError texts from the engine-backed class and the oracle."""
    from mogp_b200 import MoGP_constrained
    X = np.arange(12.).reshape(3, 4)
    Y = X[::-1] * 1.5
    for cls in (MoGP_constrained, oracle.MoGP_constrained):
        with pytest.raises(ValueError) as e:
            cls(X=X, Y=Y, **kw)
        assert msg in str(e.value)
    with pytest.raises(ValueError):
        MoGP_constrained(X=X, Y=Y[:, :3], alpha=1.)
    with pytest.raises(ValueError):
        MoGP_constrained(X=X, Y=np.ones_like(X), alpha=1.)       # zero standard deviation


def test_sweep_draw_ledger_matches_reference_order():
    """Pre-drawn sweeps consume the legacy generator exactly like the reference's interleaved per-step calls."""
    from mogp_b200 import MoGP_constrained
    X = np.arange(20.).reshape(5, 4)
    Y = np.sin(X)
    mix = MoGP_constrained(X=X, Y=Y, alpha=1., refit='sweep')
    np.random.seed(3)
    perm, a, u = mix._draws_for_sweep(5)
    np.random.seed(3)
    perm_ref = np.random.permutation(np.arange(5))
    for s in range(5):
        assert a[s] == np.random.randn(1, 1)[0, 0]
        assert u[s] == np.random.random_sample()
    assert np.array_equal(perm, perm_ref)


def test_chain_config_carries_the_sampler_options():
    """The Python facade hands the engine the reference's model constants and the look-ahead option; the struct
    layout is the C header's (include/mogp_b200.h: mogp_chain_config)."""
    import ctypes as C
    from mogp_b200 import MoGP_constrained
    X, Y = generate_toy_data(seed=0)
    mix = MoGP_constrained(X=X, Y=Y, alpha=1.5, threshold=0.5, noise_variance=0.5, refit='sweep', lookahead=24)
    cfg = mix._chain_config()
    assert (cfg.alpha, cfg.noise_variance, cfg.signal_variance, cfg.lengthscale) == (1.5, 0.5, 1.0, 4.0)
    assert (cfg.use_threshold, cfg.thr_index, cfg.threshold, cfg.fast_updates, cfg.lookahead) == (1, 1, 0.5, 1, 24)
    assert MoGP_constrained(X=X, Y=Y, alpha=1.)._chain_config().lookahead == 0      # engine default
    assert MoGP_constrained(X=X, Y=Y, alpha=1.)._chain_config().fast_updates == 0   # refit='move': reference-exact updates
    # spec (4 x int32) | 4 doubles | 2 x int32 | double | 2 x int32
    assert C.sizeof(_capi.ChainConfig) == 16 + 32 + 8 + 8 + 8
