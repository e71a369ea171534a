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


def _build_c_host(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_host")
    libdir = os.path.join(ROOT, "diaglib_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_host.c"),
                           "-o", exe, "-L" + libdir, "-ldiaglib_b200", "-Wl,-rpath," + libdir])
    return exe


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_plain_c_host_links_and_fails_loudly_without_device(tmp_path):
    """examples/c_host.c: a C program binds the drivers through include/diaglib_b200.h alone (every
    scalar by reference, Fortran argument order); without a GPU it reports the missing device"""
    import subprocess
    r = subprocess.run([_build_c_host(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_plain_c_host_solves_the_reference_test_problem(tmp_path):
    """the same C program on a B200: both drivers reproduce the dense-LAPACK eigenvalues of the
    reference's toy matrix (tests/golden/toy_dense_eigs.json)"""
    import json
    import subprocess
    r = subprocess.run([_build_c_host(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    gold = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "toy_dense_eigs.json")))["eig"])[:10]
    lines = [ln for ln in r.stdout.splitlines() if " eig:" in ln]
    assert len(lines) == 2
    for ln in lines:
        assert "ok=1 status=0" in ln
        got = np.array([float(x) for x in ln.split("eig:")[1].split()])
        assert np.abs(got - gold).max() / gold.max() < 1e-10
