"""csolve_b200 -- B200-native search hot path of the csolve constraint solver.

The product is the C-ABI shared library ``libcsolve_b200.so`` (sources under
``csolve_b200/csrc``, interface in ``include/csolve_b200.h``). This package is
the thin Python mirror used by the tests and by ``bench.py``: it loads the
library with ctypes and exposes the same entry points.
"""
from .host import (  # noqa: F401
    CsolveError,
    FlatModel,
    Comm,
    GpuGroup,
    GpuProblem,
    device_count,
    Model,
    SolveResult,
    library,
    library_path,
    OBJ_ANY, OBJ_ALL, OBJ_MIN, OBJ_MAX,
    ORDER_NONE, ORDER_SMALLEST_DOMAIN, ORDER_LARGEST_DOMAIN, ORDER_SMALLEST_VALUE, ORDER_LARGEST_VALUE,
)
from . import instances  # noqa: F401
