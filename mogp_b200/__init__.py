"""mogp_b200 — B200 (sm_100a) likelihood engine behind the MoGP / MoGP_constrained API.

Mirrors the public surface of the reference package (`mogp/__init__.py:1-2`):
`MoGP_constrained`, `utils`; plus `MoGP` (the "unconstrained" configuration) and the
boundary classes `GPobs` / `DPmix`.  All GP arithmetic runs in the CUDA library
`libmogp_b200.so` (C ABI: include/mogp_b200.h); there is no CPU fallback.
"""
import os as _os

# The look-ahead pipeline of the Gibbs sweep keeps ~25 CUDA streams busy (two scoring passes in flight, the corrections
# chained behind them, the dependent chain of the moves).  With the default 8 hardware work queues several streams share
# a queue, and a stream whose head waits for a pass blocks the urgent kernels queued behind it in the same queue for the
# whole pass (measured: 350 us per move).  Must be in the environment before the process creates its CUDA context.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .obsmodel import GPobs  # noqa: F401,E402
from .allocmodel import DPmix  # noqa: F401,E402
from .mogp_constrained import MoGP_constrained, MoGP  # noqa: F401,E402
from . import utils  # noqa: F401,E402
from . import postprocessing  # noqa: F401,E402
