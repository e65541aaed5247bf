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


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/chemps2_b200.h compiles as strict C99 and a C program can drive the library"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "chemps2_b200.h"\n'
                   'int main(void){ b2_ctx* c = 0; if (b2_ctx_create(-1, &c)) return 1; if (b2_bk_init(c, 8) != B2_ERR_STATE) return 3;\n'
                   '  b2_ctx_destroy(c); return b2_version()[0] ? 0 : 2; }\n')
    exe = tmp_path / "abi"
    lib_dir = os.path.join(ROOT, "chemps2_b200")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(exe),
                          f"-L{lib_dir}", "-lchemps2_b200", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    assert subprocess.run([str(exe)]).returncode == 0
