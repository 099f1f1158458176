"""CPU test: the __host__ __device__ field arithmetic of the product (libiop_b200/csrc/gf192.cuh, fp192.cuh)
against the oracle, incl. the worst-case carry counts of both "multiplication with holes" spacings;
and that the C-ABI library loads and exports every symbol declared in include/b200_iop.h."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_field_arithmetic_on_host():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "field_host_test.bin"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "field_host_test.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


def test_c_abi_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, "libiop_b200", "lib", "libb200iop.so")
    if not os.path.exists(lib_path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "libiop_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "b200_iop.h")).read()
    names = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 35
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in b200_iop.h but not exported"
    lib.b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.b200_version()


def test_no_cpu_fallback_without_device():
    """Without a GPU the product must fail loudly (no silent CPU path)."""
    import sys
    sys.path.insert(0, ROOT)
    import libiop_b200 as L
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    try:
        L.B200(0)
    except L.B200Error as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("B200() succeeded without a CUDA device")


def test_pinned_arena_allocator_host_logic():
    """tests/cpp/arena_test.cpp: free-list allocator behind the pinned operator new (malloc-backed slabs) and
    make_uninitialized_vector."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "arena_test.bin"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "arena_test.bin")], capture_output=True, text=True)
    assert out.returncode == 0 and "0 failed" in out.stdout, out.stdout + out.stderr
