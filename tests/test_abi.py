"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares; without
a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import diaglib_b200 as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for h in ("diaglib_b200.h", "diaglib_b200_kernels.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms |= set(re.findall(r"\b(diaglib_b200_\w+)\s*\(", txt))
    return sorted(syms)


def test_library_exports_every_declared_symbol():
    lib = D.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"


def test_header_cites_reference_lines():
    txt = open(os.path.join(ROOT, "include", "diaglib_b200.h")).read()
    for cite in ("diaglib.f90:171-172", "diaglib.f90:1483-1484", "diaglib.f90:3185", "diaglib.f90:3481",
                 "main.f90:72-90", "main.f90:146-171"):
        assert cite in txt


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_device():
    lib = D.lib()
    assert lib.diaglib_b200_init(C.c_int32(0)) == 5  # DIAGLIB_B200_ENODEVICE
    with pytest.raises(D.DiaglibError):
        D.init(0)
    ev = np.asfortranarray(np.random.default_rng(0).standard_normal((50, 4)))
    before = ev.copy()
    eig = np.zeros(4)
    ok = C.c_int32(1)
    i = lambda v: C.byref(C.c_int32(v))  # noqa: E731
    d = lambda v: C.byref(C.c_double(v))  # noqa: E731
    lib.diaglib_b200_lobpcg_driver(i(0), i(0), i(50), i(2), i(4), i(10), d(1e-8), d(0.0), None, None, None,
                                   eig.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), C.byref(ok))
    assert ok.value == 0 and np.array_equal(ev, before)
