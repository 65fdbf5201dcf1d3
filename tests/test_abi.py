"""The C-ABI library loads and exports every function include/*.h declares (no compute calls)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "**", "*.h"), recursive=True):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^[A-Za-z_][\w \t\*]*?\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M):
            if "static" in m.group(0) or "typedef" in m.group(0):
                continue
            names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    from otezip_b200.native import lib_path
    lib = ctypes.CDLL(lib_path())
    names = declared_functions()
    assert "otz_extract_run" in names and len(names) > 20
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_no_gpu_means_loud_failure_not_fallback():
    from otezip_b200.native import Lib
    import ctypes as C
    L = Lib.get().L
    if L.otz_device_count() > 0:
        return
    h = C.c_void_p()
    assert L.otz_ctx_create(0, C.byref(h)) != 0 and not h.value
    assert L.otz_last_error()
