"""chemps2_b200 — B200-native two-site DMRG sweep hot path (sigma build + operator updates) behind a C ABI.

Python here is only the host-side mirror used by tests and bench.py: it forwards to libchemps2_b200.so.
"""
from . import _lib
from ._lib import B2Error, check, lib
from .api import Context, Heff, OpSet

__all__ = ["Context", "OpSet", "Heff", "B2Error", "lib", "check"]
