"""CPU check of K1's ALGORITHM: tests/host_emul/emul.cu runs the kernel's own rule code
(bg_device.cuh, compiled for the host) through a sequential emulation of the level-by-level expansion
and must reproduce the reference's ordered legal-play lists from the golden files."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_emul", "emul.cu")
SO = os.path.join(ROOT, "tests", "host_emul", "libemul.so")


@pytest.fixture(scope="module")
def emul():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    hdr = os.path.join(ROOT, "mlp-ppo-2ply-p3_b200", "csrc", "bg_device.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-Wno-deprecated-gpu-targets",
                               "-I", os.path.dirname(hdr), "-o", SO, SRC])
    L = C.CDLL(SO)
    L.emul_legal_moves.restype = C.c_int
    L.emul_legal_moves.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return L


@pytest.mark.parametrize("name", ["initial_table", "allrolls", "adversarial"])
def test_emulated_kernel_algorithm_matches_reference(emul, name):
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    out = np.zeros((4096, 52), np.int8)
    for i in range(len(d["counts"])):
        b = np.ascontiguousarray(d["boards"][i])
        n = emul.emul_legal_moves(b.ctypes.data, int(d["players"][i]), int(d["dice"][i, 0]), int(d["dice"][i, 1]),
                                  out.ctypes.data, 4096)
        assert n == d["counts"][i], (name, i)
        lo, hi = d["offsets"][i], d["offsets"][i + 1]
        assert np.array_equal(out[:n], d["after"][lo:hi]), (name, i)
