"""mogp_b200 — B200 (sm_100a) likelihood engine behind the MoGP / MoGP_constrained API.

Mirrors the public surface of the reference package (`mogp/__init__.py:1-2`):
`MoGP_constrained`, `utils`; plus `MoGP` (the "unconstrained" configuration) and the
boundary classes `GPobs` / `DPmix`.  All GP arithmetic runs in the CUDA library
`libmogp_b200.so` (C ABI: include/mogp_b200.h); there is no CPU fallback.
"""
from .obsmodel import GPobs  # noqa: F401
from .allocmodel import DPmix  # noqa: F401
from .mogp_constrained import MoGP_constrained, MoGP  # noqa: F401
from . import utils  # noqa: F401
from . import postprocessing  # noqa: F401
