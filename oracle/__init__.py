"""CPU oracle for the MoGP likelihood path — TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the reference's algorithm
(fraenkel-lab/mogp, `mogp/obsmodel.py`, `mogp/allocmodel.py`,
`mogp/neg_linear.py`, `mogp/mogp_constrained.py`) together with the semantics
of the un-vendored third-party dependency that does the arithmetic for it:
GPy==1.10 (setup.py:14; notebooks executed on GPy 1.9.9 / paramz 0.9.5).
GPy is not installable here (no network), so its published algorithm is
restated in `oracle/gpy_restated.py`.

Parity pin: the executed tutorial notebook of the reference
(`example/tutorial_train_mogp_model.ipynb`, cells 7/10/12/16/19) — see
`tests/test_oracle_kat.py`, which checks the printed Gibbs trace, the final
assignment vector, the learned hyper-parameters, GPy's printed objective and
the two printed MAP log-likelihoods.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and only as the checker or
the timed CPU baseline.  Nothing under `mogp_b200/` imports it.
"""
from .gpy_restated import GPRegressionOracle, GammaPrior  # noqa: F401
from .reference_model import GPobs, DPmix, MoGP_constrained, MoGP  # noqa: F401
from .reference_model import generate_toy_data, rank_cluster_prediction  # noqa: F401
