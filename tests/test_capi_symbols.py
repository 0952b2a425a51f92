"""The shared library loads without a GPU and exports every symbol that include/ctdirect_b200.h declares."""
import ctypes
import os
import re


def _declared():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "include", "ctdirect_b200.h")).read()
    return sorted(set(re.findall(r"CTD_API\s+[\w\s\*]+?\b(ctd_\w+)\s*\(", txt)))


def test_header_symbols_are_exported(ctd):
    names = _declared()
    assert len(names) >= 24 and "ctd_create" in names and "ctd_eval_dev" in names
    lib = ctypes.CDLL(ctd.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_python_binding_covers_the_header(ctd):
    from ctdirect_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_c_header_compiles_as_c(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "ctdirect_b200.h"\nint main(void){ ctd_desc d; (void)d; return sizeof(ctd_handle) == sizeof(void*) ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])
