"""Weight-space evaluation of the refit objective (mogp_b200/csrc/wspace.cuh) on the CPU: the NumPy prototype
(oracle/weight_space.py) and the kernels' own phase functions - compiled for the host, tests/native/ws_host.cpp - against
the dense restatement of GPy's inference (oracle/gpy_restated.py): log marginal likelihood within 1e-11 relative, gradient
within 1e-10 of its norm, for length scales down to a ninth of the cluster's interval (WS_MAX_SPAN)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle.gpy_restated import GPRegressionOracle
from oracle.weight_space import WeightSpaceGP

HERE = os.path.dirname(os.path.abspath(__file__))


def cluster(m, seed, hi=8.0):
    rs = np.random.RandomState(seed)
    x = np.sort(rs.uniform(0, hi, m))
    x[: m // 8] = 0.0                                  # onset anchors: repeated inputs
    y = (48 - 6 * x + rs.randn(m) * 3 - 30) / 20
    return x, y


@pytest.mark.parametrize("theta", [(0.28, 1.0, 1.58, 0.0133), (0.5, 0.7, 2.4, 0.3), (0.1, 1.3, 0.95, 0.05)])
def test_numpy_prototype_matches_the_dense_gp(theta):
    a, var, ls, noise = theta
    x, y = cluster(700, 3)
    o = GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name='rbf', signal_variance=var, signal_variance_fix=False,
                           noise_variance=noise, noise_variance_fix=False, mean_func=True, lengthscale=ls, a_draw=a)
    g = WeightSpaceGP(x, y, var, ls, noise, a_slope=a, r=48)
    assert g.log_likelihood() == pytest.approx(o.log_likelihood(), rel=1e-11)
    xq = np.array([0.0, 0.4, 1.7, 3.3, 5.1, 7.9])
    yq = np.array([0.8, 0.7, 0.4, 0.0, -0.5, -1.1])
    lp_o = o.log_predictive_density(xq.reshape(-1, 1), yq.reshape(-1, 1)).ravel()
    assert np.max(np.abs(g.log_predictive_density(xq, yq) - lp_o) / np.abs(lp_o)) < 1e-9


@pytest.fixture(scope="module")
def ws_host(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("ws") / "ws_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "native", "ws_host.cpp")], check=True)
    return exe


def run_host(exe, x, y, mean_func, thetas):
    inp = "%d %d %d\n" % (len(x), int(mean_func), len(thetas))
    inp += " ".join(repr(float(v)) for v in x) + "\n" + " ".join(repr(float(v)) for v in y) + "\n"
    inp += "\n".join("%r %r %r %r" % tuple(float(v) for v in t) for t in thetas) + "\n"
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    return [np.array([float(v) for v in line.split()]) for line in out[:len(thetas)]]


@pytest.mark.parametrize("mean_func,m", [(True, 900), (False, 900), (True, 2400)])
def test_kernel_phase_functions_match_the_dense_gp(ws_host, mean_func, m):
    x, y = cluster(m, 5)
    span = x.max() - x.min()
    thetas = [(0.28, 1.0, 1.58, 0.0133), (0.5, 0.7, 2.4, 0.3), (0.3, 1.0, 4.0, 0.5), (0.2, 1.2, span / 9.0, 0.05)]
    res = run_host(ws_host, x, y, mean_func, thetas)
    for (a, var, ls, noise), v in zip(thetas, res):
        o = GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name='rbf', signal_variance=var,
                               signal_variance_fix=False, noise_variance=noise, noise_variance_fix=False, mean_func=mean_func,
                               lengthscale=ls, a_draw=a if mean_func else None)
        g = np.zeros(4)
        g[(0 if mean_func else 1):] = o._grad_lml
        assert v[5] == 0 and 10 <= v[6] <= 48          # (the kernel's numerical rank on the interval)
        assert v[0] == pytest.approx(o.log_likelihood(), rel=1e-11)
        assert np.linalg.norm(v[1:5] - g) <= 1e-10 * np.linalg.norm(g)


def test_few_distinct_inputs_are_no_problem(ws_host):
    """Nothing the data could make singular is inverted (A = s I + R S R^T): a cluster with 20 distinct inputs - fewer than
    nodes - evaluates like any other."""
    x = np.repeat(np.linspace(0, 8, 20), 20)
    y = np.sin(x) + 0.05 * np.cos(37 * np.arange(x.size))
    th = (0.2, 1.0, 2.0, 0.1)
    v = run_host(ws_host, x, y, True, [th])[0]
    o = GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name='rbf', signal_variance=th[1], signal_variance_fix=False,
                           noise_variance=th[3], noise_variance_fix=False, mean_func=True, lengthscale=th[2], a_draw=th[0])
    assert v[5] == 0
    assert v[0] == pytest.approx(o.log_likelihood(), rel=1e-11)
    assert np.linalg.norm(v[1:5] - np.asarray(o._grad_lml)) <= 1e-10 * np.linalg.norm(o._grad_lml)


def test_accuracy_over_data_shapes_and_hyper_parameters(ws_host):
    """Random clusters (uniform, crowded near the onset with a sparse tail, a monthly grid with many repeats; intervals of
    4 to 15 years) and random theta inside the region the refit uses the weight-space form for (interval <= 9 length scales;
    noise down to 1e-4, variance up to 3): LML within 1e-10 relative of the dense GP, gradient within 1e-9 of its norm -
    an order of magnitude inside the 1e-9 parity bound (worst seen over 42 such cases: 2.7e-11 / 3.7e-11)."""
    rs = np.random.RandomState(0)
    for trial in range(9):
        m = int(rs.choice([250, 450, 600]))
        hi = float(rs.choice([4.0, 8.0, 15.0]))
        kind = trial % 3
        if kind == 0:
            x = np.sort(rs.uniform(0, hi, m))
        elif kind == 1:
            x = np.sort(hi * rs.beta(0.6, 2.5, m))
        else:
            x = np.round(np.sort(rs.uniform(0, hi, m)) * 12) / 12
        x[: m // 8] = 0.0
        y = (48 - 20 * rs.uniform(0.1, 0.6) * x + rs.randn(m) * rs.uniform(0.5, 4) - 30) / 20
        span = x.max() - x.min()
        thetas = [(rs.uniform(0.0, 0.8), rs.uniform(0.2, 3.0), span / rs.uniform(1.0, 9.0), float(10 ** rs.uniform(-4, 0)))
                  for _ in range(2)]
        for (a, var, ls, noise), v in zip(thetas, run_host(ws_host, x, y, True, thetas)):
            o = GPRegressionOracle(x.reshape(-1, 1), y.reshape(-1, 1), kernel_name='rbf', signal_variance=var,
                                   signal_variance_fix=False, noise_variance=noise, noise_variance_fix=False, mean_func=True,
                                   lengthscale=ls, a_draw=a)
            g = np.asarray(o._grad_lml)
            assert v[5] == 0
            assert abs(v[0] - o.log_likelihood()) <= 1e-10 * abs(o.log_likelihood()), (trial, a, var, ls, noise)
            assert np.linalg.norm(v[1:5] - g) <= 1e-9 * np.linalg.norm(g), (trial, a, var, ls, noise)
