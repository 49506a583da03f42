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


def test_r_glue_compiles_against_stub(lib_path, tmp_path):
    """r/src/r_glue.c cannot be built with R here (no R), but it must be valid C whose calls match the prototypes of
    include/spwombling.h: compile it the way `R CMD SHLIB` would (its default flags, -pedantic, -Werror) against a
    minimal stand-in for R's headers, link it against libspwombling.so, and check that every spw_* symbol it needs is
    exported by the library and every .Call routine it registers is defined."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = os.path.join(ROOT, "r", "src", "r_glue.c")
    obj, so = str(tmp_path / "r_glue.o"), str(tmp_path / "spWombling.so")
    cmd = ["gcc", "-std=gnu11", "-I", os.path.join(ROOT, "tools", "r_stub"), "-DNDEBUG", "-I", os.path.join(ROOT, "include"),
           "-fpic", "-g", "-O2", "-Wall", "-pedantic", "-Werror", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["gcc", "-shared", "-o", so, obj, "-L", os.path.dirname(lib_path), "-lspwombling"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    und = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    needed = sorted(set(re.findall(r"\b(spw_[a-z0-9_]+)\b", und)))
    assert needed and set(needed) <= set(header_functions())
    # everything else still undefined must be R's API (resolved by libR at load time) or libc
    other = [ln.split()[-1] for ln in und.splitlines() if ln.strip() and "spw_" not in ln]
    assert all(o.startswith(("Rf_", "R_", "REAL", "INTEGER", "SET_VECTOR_ELT", "XLENGTH", "_", "__")) or "@" in o for o in other), other
    # the registration table matches the definitions and what the R side calls
    csrc = open(src).read()
    registered = re.findall(r'\{"(spw_r_[a-z_]+)", \(DL_FUNC\)&(spw_r_[a-z_]+), (\d+)\}', csrc)
    assert registered and all(a == b for a, b, _ in registered)
    rsrc = "".join(open(os.path.join(ROOT, "r", "R", f)).read() for f in sorted(os.listdir(os.path.join(ROOT, "r", "R"))))
    for name, _, nargs in registered:
        calls = re.findall(r"\.Call\(%s\b([^)]*)" % name, rsrc)
        assert calls, "%s is registered but never .Call()ed from r/R" % name
    for name in set(re.findall(r"\.Call\((spw_r_[a-z_]+)", rsrc)):
        assert name in [a for a, _, _ in registered], "%s is .Call()ed but not registered" % name


def test_r_packaging_files_exist():
    for f in ("r/src/Makevars", "r/patches/NAMESPACE.patch", "r/patches/DESCRIPTION.patch", "r/R/zzz.R", "r/R/zz_gpu_backend.R"):
        assert os.path.exists(os.path.join(ROOT, f)), f
    assert "useDynLib(spWombling, .registration = TRUE)" in open(os.path.join(ROOT, "r/patches/NAMESPACE.patch")).read()
    assert "spw_r_shutdown" in open(os.path.join(ROOT, "r/R/zzz.R")).read()
