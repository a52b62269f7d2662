"""The sequential model of the std::sort emulation (csrc/sort_emul.h: misfit-pairing partition, left-first lazy
segment walk, leaf insertion sort, heap fallback) against the real libstdc++ std::sort / __introsort_loop.
The CUDA select kernel implements exactly these steps with group-wide scans."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sortemul") / "libsortemul.so")
    subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", os.path.join(HERE, "cpu", "sort_emul_host.cpp"), "-o", out],
                   check=True)
    lib = C.CDLL(out)
    f8 = np.ctypeslib.ndpointer(np.float64, flags="C")
    i4 = np.ctypeslib.ndpointer(np.int32, flags="C")
    for fn in (lib.emul_sort_perm, lib.std_sort_perm):
        fn.argtypes = [f8, C.c_int, i4, C.c_int]
    return lib


def both(lib, keys, depth=-1):
    n = len(keys)
    a, b = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    lib.emul_sort_perm(keys, n, a, depth)
    lib.std_sort_perm(keys, n, b, depth)
    return a[:n], b[:n]


@pytest.mark.parametrize("n", [0, 1, 2, 15, 16, 17, 18, 33, 100, 1000, 5000, 40000])
@pytest.mark.parametrize("distinct", [1, 2, 3, 10, 1000, 10**6])
def test_matches_std_sort(lib, n, distinct):
    rng = np.random.default_rng(n * 7 + distinct)
    for variant in range(3):
        keys = rng.integers(0, distinct, n).astype(np.float64)
        if variant == 1:
            keys = np.sort(keys)[::-1].copy()
        if variant == 2:
            keys = np.sort(keys).copy()
        a, b = both(lib, keys)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("depth", [0, 1, 2, 5])
def test_depth_limit_exhaustion_takes_the_heap_path(lib, depth):
    rng = np.random.default_rng(depth)
    for n in (17, 100, 3000):
        keys = rng.integers(0, 50, n).astype(np.float64)
        a, b = both(lib, keys, depth)
        assert np.array_equal(a, b)


def test_port_oracle_agrees(lib, port):
    rng = np.random.default_rng(9)
    keys = rng.integers(0, 40, 5000).astype(np.float64)
    a, _ = both(lib, keys)
    assert np.array_equal(a, port.sort_perm_desc(keys))
