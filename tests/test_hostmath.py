"""Pins the FP64 numerics the CUDA kernels run (ros_stereo_slam_b200/csrc/cvmath.cuh,
fmat7.cuh) against cv2 4.13.0, using a HOST build of those same headers
(tests/hostmath/hostmath.cpp -- a test tool, never linked into libvo_b200.so).
Everything that feeds a rank-deficient null space must be bit-identical."""
import ctypes
import os
import subprocess
import sys

import cv2
import numpy as np
import pytest

from oracle import cvrng, glue, synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
DP = ctypes.POINTER(ctypes.c_double)
FP = ctypes.POINTER(ctypes.c_float)


def P(a):
    return a.ctypes.data_as(DP if a.dtype == np.float64 else FP)


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    so = os.path.join(HERE, "hostmath", "libhostmath.so")
    hdrs = [os.path.join(HERE, "..", "ros_stereo_slam_b200", "csrc", h) for h in ("cvmath.cuh", "fmat7.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-msse2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                               "-o", so, src, "-lm"])
    return ctypes.CDLL(so)


def test_jacobi_svd_bit_exact(lib):
    rng = np.random.default_rng(1)
    for n in (3, 4, 12):
        for _ in range(40):
            A = rng.standard_normal((n, n))
            if n == 12:  # rank 10, like EPnP's MtM for 5 points
                B = rng.standard_normal((10, 12))
                A = B.T @ B
            w = np.zeros(n); u = np.zeros((n, n)); vt = np.zeros((n, n))
            lib.hm_svd(P(A), n, P(w), P(u), P(vt))
            w0, u0, vt0 = cv2.SVDecomp(A)
            assert np.array_equal(w0.ravel(), w) and np.array_equal(u0, u) and np.array_equal(vt0, vt)


def test_solve_invert_multransposed_bit_exact(lib):
    rng = np.random.default_rng(2)
    for n in (3, 4, 5):
        for _ in range(40):
            A = rng.standard_normal((6, n)); b = rng.standard_normal(6); x = np.zeros(n)
            lib.hm_solve(P(A), P(b), n, P(x))
            _, x0 = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)
            assert np.array_equal(x0.ravel(), x)
    # the runtime-size pieces EPnP's beta initialisations use, serial and in the device's wavefront order (incl. rank-deficient
    # and nearly converged systems, where the speculative next sweep must leave the data alone)
    for n in (3, 4, 5):
        for k in range(60):
            A = rng.standard_normal((6, n)); b = rng.standard_normal(6)
            if k % 3 == 1:
                A[:, -1] = A[:, 0] * 2.0                      # rank deficient
            if k % 3 == 2:
                A = np.linalg.qr(rng.standard_normal((6, 6)))[0][:, :n] * rng.uniform(0.1, 10, n)   # orthogonal columns
            _, x0 = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)
            for wavefront in (0, 1):
                x = np.zeros(n)
                lib.hm_solve_rt(P(np.ascontiguousarray(A)), P(b), n, wavefront, P(x))
                assert np.array_equal(x0.ravel(), x), (n, k, wavefront)
    for _ in range(40):
        A = rng.standard_normal((3, 3)); inv = np.zeros((3, 3))
        lib.hm_invert3(P(A), P(inv))
        _, i0 = cv2.invert(A, flags=cv2.DECOMP_SVD)
        assert np.array_equal(i0, inv)
        M = rng.standard_normal((10, 12)); out = np.zeros((12, 12))
        lib.hm_mtm12(P(M), 10, 0, P(out))
        assert np.array_equal(cv2.mulTransposed(M, True), out)


@pytest.mark.parametrize("n,frac", [(5000, 0.3), (800, 0.6)])
def test_epnp5_bit_exact(lib, n, frac):
    """Every RANSAC hypothesis = cv2.solvePnP(5 pts, SOLVEPNP_EPNP), bit for bit."""
    K4 = np.array([glue.FX, glue.FY, glue.CX, glue.CY])
    X, xy, _, _, _ = synth.pnp_stress_case(n, frac, 0.3, seed=3)
    S = cvrng.sample_list(n, 5, 200)
    for idx in S:
        o = np.ascontiguousarray(X[idx]); i = np.ascontiguousarray(xy[idx])
        ok, rv, tv = cv2.solvePnP(o.reshape(-1, 1, 3), i.reshape(-1, 1, 2), glue.K, np.zeros((4, 1)),
                                  flags=cv2.SOLVEPNP_EPNP)
        r = np.zeros(3); t = np.zeros(3); R = np.zeros(9)
        lib.hm_epnp5(P(o), P(i), P(K4), 0, P(r), P(t), P(R))
        assert np.array_equal(rv.ravel(), r) and np.array_equal(tv.ravel(), t)


def test_triangulate_bit_exact(lib):
    P1, P2 = glue.projection_matrices()
    rng = np.random.default_rng(0)
    n = 5000
    z = rng.uniform(3, 80, n)
    x1 = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], 1)
    x2 = x1.copy(); x2[:, 0] -= glue.FX * 0.54 / z
    x1 = (x1 + rng.normal(0, 0.1, (n, 2))).astype(np.float32)
    x2 = (x2 + rng.normal(0, 0.1, (n, 2))).astype(np.float32)
    h0 = cv2.triangulatePoints(P1, P2, x1.T.copy(), x2.T.copy())
    xyz = np.zeros((n, 3), np.float32); h4 = np.zeros((n, 4), np.float32)
    lib.hm_triangulate(P(np.ascontiguousarray(P1)), P(np.ascontiguousarray(P2)), P(x1), P(x2), n, P(xyz), P(h4))
    assert np.array_equal(h0.T, h4)
    assert np.array_equal(glue.triangulate(P1, P2, x1, x2), xyz)


def test_fmat_7point_matches_cv2(lib):
    """Same number of models; F equal up to the last bits of cv::solveCubic's roots
    (its exact operation order is not reproduced; relative 1e-9 is far below what can
    flip an inlier)."""
    from test_oracle_ransac import _flow_case
    m1, m2 = _flow_case(2000, 2)
    S7 = cvrng.sample_list(2000, 7, 200)
    for idx in S7:
        a = np.ascontiguousarray(m1[idx]); b = np.ascontiguousarray(m2[idx])
        F0, _ = cv2.findFundamentalMat(a, b, cv2.FM_7POINT)
        F = np.zeros(27)
        n = lib.hm_fmat7(P(a), P(b), P(F))
        n0 = 0 if F0 is None else F0.shape[0] // 3
        assert n == n0
        if n:
            assert np.allclose(F0.ravel(), F[:9 * n], rtol=1e-9, atol=1e-12)
