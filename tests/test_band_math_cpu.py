"""Chunk mathematics of the banded solver (tf_band.h) on the CPU: the linear-
fractional LU chunk maps + affine substitution maps, chained by a sequential
scan with the same combine operators the GPU scans use, must reproduce a dense
LU solve."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    d = tempfile.mkdtemp(prefix="tfband_")
    so = os.path.join(d, "emul.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-I",
                           os.path.join(ROOT, "triflow_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpu_emul", "band_emul.cpp"),
                           "-o", so])
    return ctypes.CDLL(so)


def random_band(n, beta, rng, dominance):
    W = 2 * beta + 1
    A = rng.standard_normal((n, W))
    A[:, beta] = dominance * np.sum(np.abs(A), axis=1) * np.sign(A[:, beta])
    dense = np.zeros((n, n))
    for r in range(n):
        for d in range(-beta, beta + 1):
            if 0 <= r + d < n:
                dense[r, r + d] = A[r, beta + d]
    return A, dense


@pytest.mark.parametrize("beta,C", [(1, 4), (1, 8), (2, 8), (2, 4), (3, 8), (5, 8),
                                    (5, 6), (2, 2), (4, 4)])
@pytest.mark.parametrize("dominance", [1.0, 0.35])
def test_chunked_scan_solver_matches_dense(lib, beta, C, dominance):
    rng = np.random.default_rng(beta * 100 + C)
    nchunks = 37
    n = nchunks * C
    A, dense = random_band(n, beta, rng, dominance)
    f = rng.standard_normal(n)
    x = np.zeros(n)
    L = np.zeros((n, beta))
    U = np.zeros((n, beta + 1))
    P = ctypes.POINTER(ctypes.c_double)
    p = lambda a: a.ctypes.data_as(P)
    bad = lib.band_solve(beta, C, nchunks, p(A), p(f), p(x), p(L), p(U))
    assert bad == 0
    xref = np.linalg.solve(dense, f)
    assert np.max(np.abs(x - xref)) <= 1e-9 * np.max(np.abs(xref))
    # the factors are those of the sequential no-pivot LU
    Ld = np.eye(n)
    Ud = np.zeros((n, n))
    for r in range(n):
        Ud[r, r] = 1.0 / U[r, 0]
        for q in range(1, beta + 1):
            if r + q < n:
                Ud[r, r + q] = U[r, q]
            if r - q >= 0:
                Ld[r, r - q] = L[r, q - 1]
    assert np.max(np.abs(Ld @ Ud - dense)) <= 1e-9 * np.max(np.abs(dense))
