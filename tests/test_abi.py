"""The C-ABI library loads and exports every symbol include/igt_mpc.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from igt_mpc_int_b200 import _lib, build as libbuild

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    libbuild.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "igt_mpc.h")).read()
    declared = set(re.findall(r"\b(igt_[a-z_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "libigtmpc.so does not export %s" % name
    assert set(_lib.EXPORTS) <= declared


def test_params_struct_matches_header(lib):
    p = _lib.default_params(_lib.PREC_F64)
    assert (p.N, p.n_rk, p.dt) == (40, 4, 0.1)
    assert (p.v_min, p.v_max, p.a_min, p.a_max, p.df_max, p.ey_lim) == (0.0, 5.0, -4.0, 3.0, 1.0, 0.2)   # mpc.py:56-61
    assert abs(p.da_max - 0.09) < 1e-15 and abs(p.ddf_max - 0.07) < 1e-15 and p.d_min == 5.6
    assert p.precision == _lib.PREC_F64 and p.max_iter > 0 and p.second_order == 1
    p32 = _lib.default_params(_lib.PREC_F32)
    assert p32.precision == _lib.PREC_F32 and p32.tol > p.tol
    assert lib.igt_version().decode().startswith("igtmpc")


def test_argument_errors_do_not_crash(lib):
    assert lib.igt_create(None, None) == -1
    assert b"null" in lib.igt_last_error(None)
    bad = _lib.default_params()
    bad.N = 1000
    h = ctypes.c_void_p()
    assert lib.igt_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert lib.igt_launch_count(None) == 0
    lib.igt_destroy(None)
