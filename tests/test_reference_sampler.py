"""The reference's OWN sampler code (mogp/mogp_constrained.py, imported unmodified from an installed copy of the
reference) run over a substituted `GPobs` plug-in (mogp_b200/refcompat.py = INTEGRATION.md route 2).

CPU (here): over the ORACLE's GPobs it must reproduce the executed tutorial notebook - which pins the GPy restatement
under reference code rather than under another restatement - and must equal the oracle's own sampler restatement on the
paths the notebook does not exercise (threshold rule, linear kernel, no mean function): that pins `oracle.MoGP_constrained`
against `mogp.MoGP_constrained` itself.
GPU: over `mogp_b200.obsmodel.GPobs` (the CUDA engine) the reference's sampler prints the notebook's trace, and the
reference's NaN smoke test (mogp/test_package.py:30-46) runs as written.

Only k-means is patched (mogp_constrained.py:182-186): sklearn numbers the same partition differently between versions
(SURVEY.md section 8c, KAT-2), so the notebook's labels are injected."""
import logging
import re

import numpy as np
import pytest

import oracle
import oracle.reference_model as oracle_model
from conftest import tutorial_data, TUTORIAL_Z0_BLOCK
from test_oracle_kat import TRACE_1_9, Z_FINAL_10
from mogp_b200 import refcompat

needs_reference = pytest.mark.skipif(refcompat.find_reference() is None,
                                     reason="no installed copy of the reference (baseline/_ref or /root/reference)")


class _Labels:
    def __init__(self, labels):
        self.labels_ = np.asarray(labels)


class _Trace(logging.Handler):
    """Occupancies the reference logs once per sweep (mogp_constrained.py:286)."""

    def __init__(self):
        super().__init__()
        self.rows, self.init = [], None

    def emit(self, record):
        msg = record.getMessage()
        m = re.match(r"Iter (\d+): \[(.*)\]", msg, re.S)
        if m:
            self.rows.append([int(v) for v in m.group(2).split()])
        m = re.match(r"Cluster Initialization: \[(.*)\]", msg, re.S)
        if m:
            self.init = [int(v) for v in m.group(1).split()]


def run_reference(ref, X, Y, z0, tmp_path, **kw):
    ref.MoGP_constrained._k_means_init = lambda self, init_K, x_data: _Labels(np.array(z0).copy())
    h = _Trace()
    ref.logger.addHandler(h)
    try:
        mix = ref.MoGP_constrained(X=X, Y=Y, savepath=tmp_path, **kw)
        mix.sample()
    finally:
        ref.logger.removeHandler(h)
    return mix, h


@needs_reference
def test_reference_sampler_over_the_oracle_reproduces_the_notebook(tmp_path, single_blas_thread):
    ref = refcompat.load_reference_sampler(oracle_model)
    assert ref.MoGP_constrained is not oracle.MoGP_constrained
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    mix, log = run_reference(ref, X, Y, z0, tmp_path, alpha=1., num_iter=10, rand_seed=0, normalize=True)
    assert log.init == [30, 6, 12, 6, 6]                              # ipynb:L117
    assert log.rows == TRACE_1_9                                      # ipynb:L119-127
    assert list(mix.z) == Z_FINAL_10                                  # ipynb:L201-203
    m = mix.obsmodel[0].model                                         # ipynb:L272-289
    assert m[2] == pytest.approx(1.5770597538950126, rel=1e-12)
    assert m[3] == pytest.approx(0.013348237935366946, rel=1e-12)
    assert (tmp_path / "MoGP_constrained_iterinit.pkl").exists() and (tmp_path / "MoGP_constrained.pkl").exists()


@needs_reference
@pytest.mark.parametrize("kw", [dict(threshold=0.5, noise_variance=0.5), dict(kernel='linear'),
                                dict(mean_func=False, signal_variance_fix=False), dict(threshold=0.2, onset_anchor=False)],
                         ids=["threshold", "linear", "nomean-freevar", "threshold-noanchor"])
def test_oracle_sampler_restatement_equals_the_reference_sampler(kw, tmp_path, single_blas_thread):
    """Same GPobs underneath, the reference's sampler code vs the oracle's restatement of it: identical assignments, counts,
    per-patient probabilities and log-likelihood traces on the paths the notebook does not print."""
    from mogp_b200.synth import alsfrs_like, block_init
    ref = refcompat.load_reference_sampler(oracle_model)
    N = 36
    X, Y = alsfrs_like(N, seed=21)
    z0 = block_init(N, 3, X)
    common = dict(alpha=0.8, num_iter=4, rand_seed=5, normalize=True, **kw)
    mix, log = run_reference(ref, X.copy(), Y.copy(), z0, tmp_path, **common)
    orc = oracle.MoGP_constrained(X=X.copy(), Y=Y.copy(), init_z=z0, **common)
    orc.sample()
    assert list(mix.z) == list(orc.z)
    assert list(mix.allocmodel.Nk) == list(orc.allocmodel.Nk)
    assert log.rows == [list(t) for t in orc.trace]
    assert np.array_equal(mix.lobs, orc.lobs) and np.array_equal(mix.lalloc, orc.lalloc)
    for n in range(N):
        assert np.array_equal(mix.p[n], orc.p[n])
    # predict(): the reference's own code path against the restatement, same random draw
    x_new, y_new = np.array([0.25, 0.5, 1, 2.4]), np.array([46., 44., 41., 19.])
    np.random.seed(3)
    pr = mix.predict(x_new, y_new)
    np.random.seed(3)
    po = orc.predict(x_new, y_new)
    assert np.array_equal(pr, po)


@needs_reference
@pytest.mark.gpu
def test_reference_sampler_over_the_engine_reproduces_the_notebook(tmp_path):
    import mogp_b200.obsmodel as engine_model
    ref = refcompat.load_reference_sampler(engine_model)
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    mix, log = run_reference(ref, X, Y, z0, tmp_path, alpha=1., num_iter=10, rand_seed=0, normalize=True)
    assert type(mix).__module__ == "mogp.mogp_constrained"
    assert log.init == [30, 6, 12, 6, 6]                              # ipynb:L117
    assert log.rows == TRACE_1_9                                      # ipynb:L119-127
    assert list(mix.z) == Z_FINAL_10                                  # ipynb:L201-203
    assert list(np.where(mix.allocmodel.Nk > 0)[0]) == [0, 1, 3]      # ipynb:L173
    m = mix.obsmodel[0].model                                         # ipynb:L272-289
    assert m[0] == pytest.approx(0.28341388, abs=5e-8)
    assert m[2] == pytest.approx(1.5770597538950126, rel=1e-6)
    assert m[3] == pytest.approx(0.013348237935366946, rel=1e-6)
    # the checkpoint the reference writes loads back (GPModel pickles through the engine) and predicts on the data scale
    import joblib
    loaded = joblib.load(tmp_path / "MoGP_constrained.pkl")
    mu, _var = loaded.obsmodel[0].model.predict(np.array([[1.0]]))
    assert 0.0 < mu[0, 0] < 48.0
    # predict() of the reference over the engine against the engine's own facade is covered in test_chain_parity.py


@needs_reference
@pytest.mark.gpu
def test_reference_nan_smoke_test_over_the_engine(tmp_path):
    """mogp/test_package.py:30-46 as written (three patients, NaN in the middle / at the end, one initial cluster), and a
    threshold run (the -np.Inf branch, mogp_constrained.py:254-266) against the oracle's sampler."""
    import mogp_b200.obsmodel as engine_model
    ref = refcompat.load_reference_sampler(engine_model)
    X = np.vstack([[1, 2, np.nan, 4], [1, 2, 4, np.nan], [1, 2, 3, 4]]).astype(float)
    Y = np.vstack([[50, 40, np.nan, 20], [50, 40, 20, np.nan], [50, 40, 30, 20]]).astype(float)
    mix = ref.MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=5, savepath=tmp_path, rand_seed=0, normalize=True,
                               num_init_clusters=1)
    ref.MoGP_constrained._k_means_init = lambda self, init_K, x_data: _Labels(np.zeros(3, dtype=np.int32))
    mix.sample()
    assert mix.allocmodel.Nk.sum() == 3
    orc = oracle.MoGP_constrained(X=X.copy(), Y=Y.copy(), alpha=1., num_iter=5, rand_seed=0, normalize=True,
                                  num_init_clusters=1, init_z=np.zeros(3, dtype=np.int32))
    orc.sample()
    assert list(mix.z) == list(orc.z)
    from mogp_b200.synth import alsfrs_like, block_init
    N = 40
    Xs, Ys = alsfrs_like(N, seed=21)
    z0 = block_init(N, 3, Xs)
    kw = dict(alpha=0.8, num_iter=4, rand_seed=5, normalize=True, threshold=0.5, noise_variance=0.5)
    mix2, log = run_reference(ref, Xs.copy(), Ys.copy(), z0, tmp_path, **kw)
    orc2 = oracle.MoGP_constrained(X=Xs.copy(), Y=Ys.copy(), init_z=z0, **kw)
    orc2.sample()
    assert log.rows == [list(t) for t in orc2.trace]
    assert list(mix2.z) == list(orc2.z)
