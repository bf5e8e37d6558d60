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


def test_small_point_sets_follow_opencv():
    """Below 15 points cv2.findFundamentalMat(FM_RANSAC) switches estimator: N == 7 -> 7-point result, mask of ones;
    8..14 -> LMedS (restated in replay.fmat_lmeds); N < 7 -> nothing.  solvePnPRansac with 5 / 4 points is one
    EPnP / P3P solve with every point an inlier."""
    import cv2
    from oracle import synth
    sc = synth.Scene(0)
    L0, L1 = sc.render(0, "L"), sc.render(1, "L")
    pts = glue.dense_keypoint_extractor(376, 1241, 30)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(L0, L1, pts.reshape(-1, 1, 2), None)
    ok = st.ravel() == 1
    a, b = pts[ok], p1.reshape(-1, 2)[ok]
    rng = np.random.default_rng(0)
    same = total = 0
    for n in range(8, 15):
        for trial in range(4):
            sel = rng.choice(len(a), n, replace=False)
            x, y = a[sel].copy(), b[sel].copy()
            if trial % 2:
                y[:2] += rng.normal(0, 5, (2, 2)).astype(np.float32)
            F, mask = cv2.findFundamentalMat(x, y, cv2.FM_RANSAC, 1.0, 0.99)
            F2, m2 = replay.fmat_lmeds(x, y, 0.99)
            total += 1
            same += int(mask is not None and np.array_equal(mask.ravel(), m2))
            if n == 14:     # the only size whose median involves a point outside the sample: not noise-decided
                assert mask is not None and np.array_equal(mask.ravel(), m2)
    assert same >= 0.85 * total, (same, total)
    s7 = rng.choice(len(a), 7, replace=False)     # (the first grid points are collinear: OpenCV asserts on those)
    F, mask = cv2.findFundamentalMat(a[s7], b[s7], cv2.FM_RANSAC, 1.0, 0.99)
    assert F.shape[0] in (3, 6, 9) and np.all(mask.ravel() == 1)
    F, mask = cv2.findFundamentalMat(a[s7[:6]], b[s7[:6]], cv2.FM_RANSAC, 1.0, 0.99)
    assert F is None and mask is None
    X, xy, _, _, _ = synth.pnp_stress_case(50, 0.0, 0.2, seed=1)
    for n, flag in ((5, cv2.SOLVEPNP_EPNP), (4, cv2.SOLVEPNP_P3P)):
        ok_, r, t, inl = cv2.solvePnPRansac(X[:n].reshape(-1, 1, 3), xy[:n].reshape(-1, 1, 2), glue.K, np.zeros((4, 1)), None,
                                            None, False, 100, 1.0, 0.99)
        ok2, r2, t2 = cv2.solvePnP(X[:n].reshape(-1, 1, 3), xy[:n].reshape(-1, 1, 2), glue.K, np.zeros((4, 1)), flags=flag)
        assert ok_ and np.array_equal(inl.ravel(), np.arange(n)) and np.array_equal(r, r2) and np.array_equal(t, t2)
