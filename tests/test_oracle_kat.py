"""Pin the CPU oracle to the reference's only golden outputs: the executed tutorial notebook
/root/reference/example/tutorial_train_mogp_model.ipynb (cells 7, 10, 12, 16, 19).
Values below are transcribed from the notebook's stored outputs (raw-JSON lines cited)."""
import numpy as np
import pytest

import oracle
from conftest import tutorial_data, TUTORIAL_Z0_BLOCK

# ipynb:L117-127 (num_iter=10) and :L367-391 (num_iter=30, sweeps 10..29)
TRACE_1_9 = [[41, 2, 3, 14], [33, 3, 1, 23], [23, 15, 22], [21, 16, 23], [19, 20, 21], [20, 19, 21],
             [21, 19, 20], [20, 19, 21], [21, 19, 20]]
TRACE_10_29 = [[20, 19, 21], [20, 19, 21], [21, 18, 21], [20, 19, 21], [20, 19, 21], [20, 19, 21], [19, 20, 21],
               [20, 19, 21], [20, 20, 20], [20, 19, 21], [20, 19, 21], [20, 19, 21], [21, 18, 21], [20, 19, 21],
               [21, 18, 21], [20, 19, 21], [20, 19, 21], [21, 19, 20], [20, 19, 21], [21, 19, 20]]
# ipynb:L201-203
Z_FINAL_10 = [3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 0, 3, 3, 0, 3, 3, 3, 3, 3, 0, 0,
              0, 0, 1, 0, 0, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 3, 1, 1, 1, 1,
              1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]


def _chain(num_iter):
    X, Y = tutorial_data()
    z0 = np.tile(np.array(TUTORIAL_Z0_BLOCK), 3)
    mix = oracle.MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=num_iter, rand_seed=0, normalize=True, init_z=z0)
    mix.sample()
    return mix


@pytest.fixture(scope="module")
def chain10(single_blas_thread):
    return _chain(10)


def test_kat1_normaliser(chain10):
    assert chain10.mean == pytest.approx(29.840974308274728, rel=1e-15)
    assert chain10.std == pytest.approx(20.681274134508136, rel=1e-15)


def test_kat2_init(chain10):
    assert list(chain10.init_occupancy) == [30, 6, 12, 6, 6]          # ipynb:L117


def test_kat3_trace(chain10):
    assert [list(t) for t in chain10.trace] == TRACE_1_9


def test_kat4_assignments(chain10):
    assert list(chain10.z) == Z_FINAL_10
    assert list(np.where(chain10.allocmodel.Nk > 0)[0]) == [0, 1, 3]  # ipynb:L173


def test_kat5_hyperparameters(chain10):
    m = chain10.obsmodel[0].model                                     # ipynb:L272-289
    assert m.param_names == ['neg_linmap.A', 'rbf.variance', 'rbf.lengthscale', 'Gaussian_noise.variance']
    assert m[0] == pytest.approx(0.28341388, abs=5e-9)
    assert m[1] == 1.0
    assert m[2] == pytest.approx(1.5770597538950126, rel=1e-12)
    assert m[3] == pytest.approx(0.013348237935366946, rel=1e-12)
    assert m.objective() == pytest.approx(-37.289086611673795, rel=1e-13)
    # Ga(2.2, 3.3), Ga(1.8, 0.44), Ga(9, 12) as printed
    pri = [(round(p.a, 1), round(p.b, 2)) for p in m.priors if p is not None]
    assert pri == [(2.2, 3.33), (1.8, 0.44), (9.0, 12.0)]


@pytest.mark.slow
def test_kat3_kat6_thirty_iterations(single_blas_thread):
    mix = _chain(30)
    assert [list(t) for t in mix.trace] == TRACE_1_9 + TRACE_10_29
    ll = mix.lobs + mix.lalloc
    assert ll[26] == pytest.approx(77.70252573010355, rel=1e-12)      # ipynb:L385
    assert ll[28] == pytest.approx(77.70252782040932, rel=1e-12)      # ipynb:L389
    assert mix.best_ll == ll[28]
