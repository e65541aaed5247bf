"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "chemps2_b200", "libchemps2_b200.so"))
    syms = declared_symbols()
    assert len(syms) > 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_header():
    from chemps2_b200._lib import SIGNATURES
    bound = {n for n, _, _ in SIGNATURES}
    assert set(declared_symbols()) <= bound, sorted(set(declared_symbols()) - bound)


def test_version_and_error_strings():
    from chemps2_b200._lib import lib
    assert b"sm_100a" in lib.b2_version()
    import ctypes as C
    out = C.c_void_p()
    assert lib.b2_ctx_create(-1, C.byref(out)) == 0
    assert lib.b2_bk_init(out, 10) != 0      # no problem set yet -> error code + message, never a crash
    assert b"problem" in lib.b2_last_error()
    lib.b2_ctx_destroy(out)
