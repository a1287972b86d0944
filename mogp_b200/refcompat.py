"""INTEGRATION.md route 2 as code: run the reference's OWN, unmodified sampler (`mogp/mogp_constrained.py`) with its
`GPobs` plug-in replaced.

The reference binds its obsmodel with one import line, `from mogp.obsmodel import GPobs`
(mogp/mogp_constrained.py:15), and needs GPy for exactly one more thing at that level,
`GPy.util.normalizer.Standardize` (mogp/mogp_constrained.py:17, 155-158).  `load_reference_sampler` imports the
reference package from an installed, unmodified copy with

  * `mogp.obsmodel` pre-registered as the given module (default: `mogp_b200.obsmodel`, the CUDA engine), so the import
    at mogp_constrained.py:15 binds the replacement and the reference's `obsmodel.py` / `neg_linear.py` (the only
    modules that import GPy's models) are never executed;
  * a three-line stand-in for `GPy.util.normalizer` when GPy is not installed;
  * an empty `matplotlib` stand-in when matplotlib is not installed (`mogp/utils.py:2` imports it at module level; only
    `plot_mogp` uses it);
  * `numpy.Inf` restored when the installed NumPy (>= 2.0) has dropped it (mogp_constrained.py:263 spells -inf that way).

Nothing in the reference's files is edited or copied: `MoGP_constrained.sample()` / `.predict()` run as written, every
`GPobs(...)`, `calc_score`, `get_pred_at_loc`, `update_data`, `update_suffstats`, `calc_ll` call landing on the replacement.
Where the reference lives: `pip install --no-deps --target baseline/_ref /root/reference` (what `__graft_entry__.build()`
does when the reference is mounted) or any directory that holds its `mogp/` package.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_DIRS = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def find_reference(path=None):
    for d in ([path] if path else DEFAULT_DIRS):
        if d and os.path.isfile(os.path.join(d, "mogp", "mogp_constrained.py")):
            return d
    return None


def _stub_gpy(standardize_cls):
    try:
        import GPy  # noqa: F401  (a real install wins)
        return
    except ImportError:
        pass
    gpy = types.ModuleType("GPy")
    util = types.ModuleType("GPy.util")
    norm = types.ModuleType("GPy.util.normalizer")
    norm.Standardize = standardize_cls
    util.normalizer = norm
    gpy.util = util
    gpy.__stub__ = True
    sys.modules.update({"GPy": gpy, "GPy.util": util, "GPy.util.normalizer": norm})


def load_reference_sampler(obsmodel_module=None, path=None):
    """Import the reference's `mogp` package with `mogp.obsmodel` := obsmodel_module; returns the reference's
    `mogp.mogp_constrained` module (its `MoGP_constrained` class is the reference's own code)."""
    ref_dir = find_reference(path)
    if ref_dir is None:
        raise ImportError("reference package not found (looked in {})".format(", ".join(DEFAULT_DIRS)))
    if obsmodel_module is None:
        from . import obsmodel as obsmodel_module
    import numpy as np
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                                            # mogp_constrained.py:263 on NumPy >= 2.0
    _stub_gpy(obsmodel_module.Standardize)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        sys.modules["matplotlib"] = types.ModuleType("matplotlib")
    # a fresh import each time: the replacement is bound at import (mogp_constrained.py:15)
    for name in [k for k in sys.modules if k == "mogp" or k.startswith("mogp.")]:
        del sys.modules[name]
    sys.modules["mogp.obsmodel"] = obsmodel_module
    sys.path.insert(0, ref_dir)
    try:
        mod = importlib.import_module("mogp.mogp_constrained")
    finally:
        sys.path.remove(ref_dir)
    if os.path.realpath(os.path.dirname(mod.__file__)) != os.path.realpath(os.path.join(ref_dir, "mogp")):
        raise ImportError("imported {} instead of the reference under {}".format(mod.__file__, ref_dir))
    if mod.GPobs is not obsmodel_module.GPobs:
        raise ImportError("the reference did not bind the replacement GPobs")
    return mod
