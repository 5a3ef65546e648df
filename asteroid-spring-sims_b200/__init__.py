"""asteroid-spring-sims_b200: B200-native (sm_100a) per-step force + integration hot path of the REBOUND-based
mass-spring asteroid simulator, behind the C-ABI of include/ass_b200.h.

csrc/   CUDA kernels and the extern "C" layer  -> libass_b200.so
host/   C++ mirror of the reference's own symbols (reb_integrate, reb_step, spring_forces, ...) over that C-ABI
api.py  ctypes binding used by tests/ and bench.py (plumbing; not a second implementation)
launcher.py  host-side multi-GPU launch logic (member dealing, slab bounds, the sharded step protocol) over torch.distributed

The directory name carries a hyphen, so it is imported through the sibling shim package
``asteroid_spring_sims_b200``.
"""
from . import build  # noqa: F401
from . import api  # noqa: F401
from . import launcher  # noqa: F401
from .api import *  # noqa: F401,F403
