"""SURVEY.md §8(f) rows 2-3 on the engine: monotonicity check / cluster splitting
(analysis/model_postprocessing.py:6-104) and rank-then-forecast errors (analysis/analysis_utils.py:80-92,493-521)
against the oracle's loop-for-loop restatement, on the same fitted tutorial model."""
import numpy as np
import pytest

import oracle
from oracle import postprocessing as opp
from conftest import tutorial_data, TUTORIAL_Z0_BLOCK

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fitted():
    from mogp_b200 import _capi, MoGP_constrained
    eng = _capi.Engine(0)
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    kw = dict(alpha=1., num_iter=6, rand_seed=0, normalize=True, init_z=z0)
    orc = oracle.MoGP_constrained(X=X, Y=Y, **kw)
    orc.sample()
    mix = MoGP_constrained(X=X, Y=Y, engine=eng, **kw)
    mix.sample()
    assert list(mix.z) == list(orc.z)
    yield mix, orc, X, Y
    eng.close()


def test_cluster_jumps_and_monotonicity(fitted):
    from mogp_b200 import postprocessing as pp
    mix, orc, X, Y = fitted
    rises = []
    for k in np.where(orc.allocmodel.Nk > 0)[0]:
        xs = np.arange(np.min(orc.obsmodel[k].X), np.max(orc.obsmodel[k].X) - 1, 0.1)
        m = orc.obsmodel[k].model.predict(np.r_[xs, xs + 1].reshape(-1, 1))[0].ravel()
        rises.append(np.max(m[xs.size:] - m[:xs.size]))
    for thresh in (10.0, float(np.max(rises)) + 0.5, float(np.max(rises)) - 0.5, float(np.min(rises)) - 0.5):
        for k in np.where(orc.allocmodel.Nk > 0)[0]:
            assert pp.find_cluster_jumps(mix.obsmodel[k], thresh=thresh) == opp.find_cluster_jumps(orc.obsmodel[k], thresh=thresh)
        assert pp.check_model_monotonicity(mix, thresh=thresh) == opp.check_model_monotonicity(orc, thresh=thresh)
        assert pp.check_model_monotonicity(mix, thresh=thresh, num_obs=1) == opp.check_model_monotonicity(orc, thresh=thresh, num_obs=1)


def test_rank_and_forecast_errors(fitted):
    from mogp_b200 import postprocessing as pp
    mix, orc, X, Y = fitted
    rs = np.random.RandomState(4)
    XA, YA = X[::7].copy(), Y[::7].copy() + rs.randn(*Y[::7].shape)
    XA[1, 3:], YA[1, 3:] = np.nan, np.nan                       # a ragged query row
    np.random.seed(21)
    top_o, err_o = opp.rank_and_forecast_errors(orc, XA, YA)
    np.random.seed(21)
    top_e, err_e = pp.rank_and_forecast_errors(mix, XA, YA)
    assert list(top_e) == list(top_o)
    assert np.allclose(err_e, err_o, rtol=1e-6)
    k = int(top_o[0])
    xr = XA[0][~np.isnan(XA[0])]
    assert np.allclose(pp.calc_y_pred_model(mix, xr, k), orc.obsmodel[k].model.predict(xr.reshape(-1, 1))[0].T, rtol=1e-6)


def test_split_nonmonotonic_clusters(fitted):
    """Force splits with a negative threshold: both sides retire the same ids, build the same two halves with k-means
    and refit them to the same optimum."""
    from mogp_b200 import postprocessing as pp
    mix, orc, X, Y = fitted
    n_before = len(orc.allocmodel.Nk)
    np.random.seed(5)
    opp.split_nonmonotonic_clusters(orc, rand_seed=0, thresh=-1e9)
    np.random.seed(5)
    pp.split_nonmonotonic_clusters(mix, rand_seed=0, thresh=-1e9)
    assert len(orc.allocmodel.Nk) > n_before
    assert list(mix.allocmodel.Nk) == list(orc.allocmodel.Nk)
    assert list(mix.z) == list(orc.z)
    for k in np.where(orc.allocmodel.Nk > 0)[0]:
        assert np.allclose(mix.obsmodel[k].model.param_array, orc.obsmodel[k].model.param_array, rtol=1e-4)
        assert mix.obsmodel[k].model.normalizer is not None
        xs = np.linspace(0, 6, 5).reshape(-1, 1)
        assert np.allclose(mix.obsmodel[k].model.predict(xs)[0], orc.obsmodel[k].model.predict(xs)[0], rtol=1e-5, atol=1e-6)
