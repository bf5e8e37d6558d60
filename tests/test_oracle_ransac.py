"""Pins oracle.replay (explicit RANSAC loops) bit-identical to the OpenCV entry
points the reference calls (tracking.cpp:34,75; keyFrameManagement.cpp:84,88)."""
import numpy as np
import cv2
import pytest

from oracle import replay, synth, glue, cvrng


def _flow_case(n, seed, outlier_frac=0.2, grid=False):
    """Synthetic two-view correspondences of a non-planar scene."""
    rng = np.random.default_rng(seed)
    if grid:
        g = glue.dense_keypoint_extractor(376, 1241, 9)
        sel = np.sort(rng.permutation(len(g))[:n])
        x1 = g[sel].astype(np.float64)
        n = len(x1)
    else:
        x1 = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], 1)
    z = rng.uniform(4, 60, n)
    X = np.stack([(x1[:, 0] - glue.CX) / glue.FX * z, (x1[:, 1] - glue.CY) / glue.FY * z, z], 1)
    rvec = np.array([0.002, 0.007, -0.001])
    tvec = np.array([0.02, -0.01, -0.85])
    p, _ = cv2.projectPoints(X.reshape(-1, 1, 3), rvec, tvec, glue.K, np.zeros((4, 1)))
    x2 = p.reshape(-1, 2) + rng.normal(0, 0.15, (n, 2))
    k = int(n * outlier_frac)
    o = rng.permutation(n)[:k]
    x2[o] += rng.uniform(-30, 30, (k, 2))
    return x1.astype(np.float32), x2.astype(np.float32)


def test_rng_first_samples():
    # SURVEY.md appendix A.3 regression values
    assert list(cvrng.sample_list(500, 5, 1)[0]) == [105, 4, 440, 173, 331]
    assert list(cvrng.sample_list(5000, 5, 1)[0]) == [3605, 4004, 3940, 2173, 1831]
    assert list(cvrng.sample_list(20000, 5, 1)[0]) == [3605, 19004, 8940, 7173, 1831]


@pytest.mark.parametrize("n,seed,thr,grid", [(350, 0, 1.0, True), (350, 1, 3.0, True), (2000, 2, 1.0, False),
                                             (5000, 3, 3.0, False), (40, 4, 1.0, False), (15, 5, 3.0, False)])
def test_fmat_ransac_bit_identical(n, seed, thr, grid):
    m1, m2 = _flow_case(n, seed, grid=grid)
    F0, mask0 = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, thr, 0.99)
    r = replay.fmat_ransac(m1, m2, thr, 0.99)
    assert np.array_equal(mask0.ravel(), r["mask"])
    assert np.array_equal(F0, r["F"])
    # replaying the recorded sample list reproduces the same result
    r2 = replay.fmat_ransac(m1, m2, thr, 0.99, samples=r["samples"])
    assert np.array_equal(r2["mask"], r["mask"]) and r2["best"] == r["best"]


def test_fmat_draw_samples_matches_loop():
    m1, m2 = _flow_case(350, 7, grid=True)
    r = replay.fmat_ransac(m1, m2, 1.0, 0.99)
    s, _rej = replay.fmat_draw_samples(m1, m2, len(r["samples"]))
    assert np.array_equal(s, r["samples"])


@pytest.mark.parametrize("n,frac,iters,thr,conf", [(500, 0.1, 100, 1.0, 0.99), (5000, 0.3, 100, 1.0, 0.99),
                                                  (20000, 0.5, 100, 1.0, 0.99), (3000, 0.5, 400, 8.0, 0.98)])
def test_pnp_ransac_bit_identical(n, frac, iters, thr, conf):
    X, xy, _, _, _ = synth.pnp_stress_case(n, frac, 0.3, seed=3)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X.reshape(-1, 1, 3), xy.reshape(-1, 1, 2), glue.K, np.zeros((4, 1)),
                                             None, None, False, iters, thr, conf)
    r = replay.pnp_ransac(X, xy, glue.K, iters, thr, conf)
    assert ok and r["ok"]
    assert np.array_equal(inl.ravel(), r["inliers"])
    assert np.array_equal(rvec.ravel(), r["rvec"])
    assert np.array_equal(tvec.ravel(), r["tvec"])
    r2 = replay.pnp_ransac(X, xy, glue.K, iters, thr, conf, samples=r["samples"])
    assert np.array_equal(r2["inliers"], r["inliers"]) and np.array_equal(r2["rvec"], r["rvec"])
    # the default sample list is OpenCV's RNG stream
    assert np.array_equal(cvrng.sample_list(n, 5, len(r["samples"])), r["samples"])
