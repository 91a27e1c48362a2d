"""The C-ABI library loads without a GPU and exports every symbol include/lps_b200.h declares; the
ctypes table in _native.py covers exactly those symbols; the product package never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lps_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lps_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from linear_program_solver_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _native.load()


def test_header_symbols_are_exported_and_bound(lib):
    from linear_program_solver_b200 import _native
    names = declared_functions()
    assert len(names) >= 30
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} declared in lps_b200.h but not exported"
    assert sorted(_native.SIGNATURES) == names


def test_library_answers_without_a_gpu(lib):
    assert lib.lps_version() >= 100
    assert lib.lps_grid_ld(4000) % 16 == 0 and lib.lps_grid_ld(4000) >= 4001
    assert lib.lps_exact_ld(4000) % 32 == 0
    assert lib.lps_cand_bytes(100) == 64 + 8 * lib.lps_grid_ld(100)


def test_no_cpu_fallback_and_no_oracle_import_in_the_product():
    pkg = os.path.join(ROOT, "linear_program_solver_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    code = ("import torch, numpy as np\n"
            "from linear_program_solver_b200 import Tableau, NativeUnavailable\n"
            "t = Tableau(1, 2, mode='fp64'); t.setA([[1, 1]])\n"
            "try:\n    t.pivot(0, 0)\nexcept NativeUnavailable as e:\n    print('LOUD', e)\n")
    if __import__("torch").cuda.is_available():
        pytest.skip("a GPU is present: the device path works here")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert "LOUD" in out.stdout, out.stdout + out.stderr


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/lps_b200.h must compile as C99 (no C++-isms, no torch types) and a C caller must
    link against the shared library and reach it.  Only entry points that need no GPU are called."""
    from linear_program_solver_b200 import _native
    src = tmp_path / "caller.c"
    src.write_text(
        '#include <stdio.h>\n#include "lps_b200.h"\n'
        "int main(void) {\n"
        "    lps_handle *h = NULL;\n"
        "    int rc = lps_pivot(h, 0, 0);                 /* a null handle is refused, not dereferenced */\n"
        '    printf("%d %d %d %s\\n", lps_version(), (int)lps_grid_ld(4000), rc < 0, lps_last_error() ? "msg" : "null");\n'
        "    return 0;\n}\n")
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_native.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
           "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(_native.LIB_PATH), "-Wl,-rpath," + libdir]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    version, ld, refused, msg = out.stdout.split()
    assert int(version) >= 100 and int(ld) % 16 == 0 and refused == "1" and msg == "msg"
