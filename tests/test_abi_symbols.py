"""The C-ABI libraries load on a CPU-only host and export every symbol include/*.h declares.
No compute call is made here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "thirring2d_b200")


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(names))


def test_handle_abi_exports_every_declared_symbol():
    names = declared_functions("thirring_b200.h")
    assert "tb_create" in names and "tb_cg" in names and len(names) >= 24
    lib = ctypes.CDLL(os.path.join(PKG, "libthirring_b200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_mirror_types_every_declared_symbol():
    from thirring2d_b200.lib import SIGNATURES, load_library

    load_library()
    assert sorted(SIGNATURES) == declared_functions("thirring_b200.h")


def test_reference_signature_shim_exports_family_A():
    names = declared_functions("thirring_hmc_abi.h")
    for n in ("fm_mul", "fm_conjugate_mul", "fmdm_invert_cg", "fm_invert_cg", "fmdm_mul", "alloc_vector",
              "free_vector", "test_conjugate", "tb_hmc_configure"):
        assert n in names
    lib = ctypes.CDLL(os.path.join(PKG, "libthirring_hmc.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_coarse_override_library_exports_update_gauge():
    lib = ctypes.CDLL(os.path.join(PKG, "libthirring_hmc_coarse.so"))
    assert hasattr(lib, "update_gauge")
    assert os.path.exists(os.path.join(PKG, "libthirring_b200.a"))   # the static variant of the handle library


def test_vec_ops_replacement_exports_family_B():
    names = declared_functions("thirring_vecops_abi.h")
    for n in ("fM", "fM_transpose", "cg_MdM", "cg_propagator", "vec_dot", "vec_dmul_add", "alloc_vector", "free_vector",
              "vec_zero", "vec_one", "vec_add", "vec_d_mul", "vec_zero_occupied", "tb_vecops_configure",
              "cg_MdM_occupied", "fM_occupied", "fM_occupied_sq", "action", "vec_gaussian", "alloc_field"):
        assert n in names
    lib = ctypes.CDLL(os.path.join(PKG, "libthirring_vecops.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    """Creating a context on a host without a GPU must fail loudly, never fall back."""
    import thirring2d_b200 as tb

    lib = tb.load_library()
    if lib.tb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(tb.TBError, match="no CUDA device"):
        tb.Context(8, 8)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under thirring2d_b200/ or include/ may reference it."""
    bad = []
    for base in (PKG, os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(base):
            if "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".c", ".cu", ".cuh", ".h", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"\boracle\b|pyoracle|liboracle|_ref/", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
