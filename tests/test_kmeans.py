"""k-means initialisation on the device (mogp/mogp_constrained.py:182-207) against
scikit-learn's KMeans on the same data and seed: identical labels (the seeding consumes the same RandomState draws in the
same order and numbers the centres the same way), hence the same initial partition for the sampler."""
import numpy as np
import pytest

from conftest import tutorial_data, TUTORIAL_Z0_BLOCK

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from mogp_b200 import _capi
    e = _capi.Engine(0)
    yield e
    e.close()


def same_partition(a, b):
    pairs = set(zip(a.tolist(), b.tolist()))
    return len(pairs) == len(set(a.tolist())) == len(set(b.tolist()))


def first_visit_times(N, seed):
    from mogp_b200.synth import alsfrs_like
    X, _Y = alsfrs_like(N, seed=seed)
    return X[:, 1].copy()


@pytest.mark.parametrize("N,K,seed", [(200, 4, 0), (123, 7, 5), (500, 10, 7), (2000, 40, 0), (3000, 60, 1), (10000, 30, 0),
                                      (10000, 200, 3)])
def test_device_kmeans_equals_sklearn(eng, N, K, seed):
    from sklearn.cluster import KMeans
    x = first_visit_times(N, seed + 1)
    km = KMeans(n_clusters=K, random_state=seed, n_init=10).fit(x.reshape(-1, 1))
    labels, info = eng.kmeans_init(x, K, seed, n_init=10, details=True)
    assert same_partition(labels, km.labels_), "partition differs from sklearn's"
    assert np.array_equal(labels, km.labels_), "same partition but numbered differently"
    assert info['inertia'] == pytest.approx(km.inertia_, rel=1e-9)
    assert np.allclose(np.sort(info['centers']), np.sort(km.cluster_centers_.ravel()), rtol=1e-9, atol=1e-12)


def test_tutorial_initialisation(eng):
    """The reference notebook's 'Cluster Initialization: [30 6 12 6 6]' (ipynb:L117): the device k-means finds that
    partition (scikit-learn 0.21 numbered two of its clusters the other way round), and the
    sampler started without init_z uses it."""
    from sklearn.cluster import KMeans
    from mogp_b200 import MoGP_constrained
    X, Y = tutorial_data()
    labels = eng.kmeans_init(X[:, 1], 5, 0, n_init=10)
    assert same_partition(labels, np.tile(np.array(TUTORIAL_Z0_BLOCK), 3))
    assert sorted(np.bincount(labels).tolist()) == [6, 6, 6, 12, 30]
    assert np.array_equal(labels, KMeans(n_clusters=5, random_state=0, n_init=10).fit(X[:, 1].reshape(-1, 1)).labels_)
    mix = MoGP_constrained(X=X, Y=Y, alpha=1., num_iter=2, rand_seed=0, normalize=True, engine=eng)
    mix.initialize_sampler(5, True)
    assert np.array_equal(mix.z, labels)


def test_duplicates_and_small_inputs(eng):
    """Repeated values (the tutorial repeats every time three times), K = 1, K = n."""
    from sklearn.cluster import KMeans
    rs = np.random.RandomState(4)
    x = np.repeat(rs.uniform(0, 5, 25), 4)
    rs.shuffle(x)
    for K in (1, 3, 8):
        km = KMeans(n_clusters=K, random_state=2, n_init=10).fit(x.reshape(-1, 1))
        assert np.array_equal(eng.kmeans_init(x, K, 2), km.labels_)
    y = np.array([0.3, 2.0, 2.1, 7.5])
    assert same_partition(eng.kmeans_init(y, 4, 0), np.arange(4))
    with pytest.raises(Exception):
        eng.kmeans_init(y, 5, 0)
