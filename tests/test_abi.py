"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol declared in
include/spwombling.h, and the ctypes table covers exactly that set (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spwombling.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spw_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from spwombling_b200.build import build
    return build()


def test_header_declares_entry_points():
    names = header_functions()
    for must in ("spw_cov", "spw_cross_cov", "spw_hlm_run", "spw_gradient_run", "spw_hpd_summary",
                 "spw_gradient_shard_dev", "spw_last_error", "spw_shutdown"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in header_functions():
        assert hasattr(lib, name), "libspwombling.so does not export %s" % name


def test_ctypes_table_matches_header(lib_path):
    from spwombling_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    _lib.load()
    assert _lib.load().spw_version() >= 100


def test_no_product_import_of_oracle():
    """The product package must never import the CPU oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "spwombling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_argument_errors_without_gpu(lib_path):
    from spwombling_b200 import api
    with pytest.raises(ValueError):
        api.cov_matrix([[0.0, 0.0]], "exponential", 1.0)      # only the three named kernels
    with pytest.raises(ValueError):
        api.hlmBayes_sp(y=[1.0, 2.0], coords=[[0, 0], [1, 1]], niter=10, cov_type="matern1", verbose=False)


def test_r_glue_compiles_against_stub():
    """r/src/r_glue.c cannot be built here (no R), but it must at least be valid C whose calls match the prototypes
    of include/spwombling.h: compile it with -fsyntax-only against a minimal stand-in for R's headers."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    cmd = ["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tools", "r_stub"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "r_glue.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
