"""Replayable RANSAC loops.  TEST INFRASTRUCTURE ONLY.

Explicit restatements of the two OpenCV RANSAC drivers the reference calls
(cv::findFundamentalMat at src/tracking.cpp:34,75 and cv::solvePnPRansac at
src/keyFrameManagement.cpp:84,88 of the reference tree).  The outer loop
(RANSACPointSetRegistrator::run: sampling, record-setter rule, adaptive
iteration count) is restated; the minimal solvers are OpenCV's own
(cv2.findFundamentalMat(FM_7POINT), cv2.solvePnP(SOLVEPNP_EPNP)) so every
hypothesis is the reference's hypothesis.  tests/test_oracle_ransac.py pins the
loops bit-identical to cv2.findFundamentalMat / cv2.solvePnPRansac.

Each loop can be driven by OpenCV's RNG (default) or by an explicit list of
minimal-sample index tuples (``samples``), and records what it did so the CUDA
path can be compared hypothesis by hypothesis.
"""
import numpy as np
import cv2

from .cvrng import CvRNG, draw_subset, ransac_update_num_iters, RANSAC_SEED

FLT_EPSILON = float(np.finfo(np.float32).eps)


# ----------------------------------------------------------------------------
# fundamental matrix
# ----------------------------------------------------------------------------
def have_collinear_points(pts, count):
    """haveCollinearPoints (calib3d fundam.cpp): is the LAST of ``count`` points
    on a line through two earlier ones (or too close)?  pts float32 (>=count,2)."""
    i = count - 1
    p = pts.astype(np.float64)
    for j in range(i):
        dx1 = p[j, 0] - p[i, 0]
        dy1 = p[j, 1] - p[i, 1]
        for k in range(j):
            dx2 = p[k, 0] - p[i, 0]
            dy2 = p[k, 1] - p[i, 1]
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def fmat_check_subset(m1, m2, idx):
    s1 = m1[idx]
    s2 = m2[idx]
    return not have_collinear_points(s1, len(idx)) and not have_collinear_points(s2, len(idx))


def fmat_draw_samples(m1, m2, n_samples, seed=RANSAC_SEED, max_attempts=10000):
    """The first n_samples ACCEPTED 7-tuples RANSACPointSetRegistrator::getSubset
    would return for (m1, m2), including collinearity rejections.  Returns
    (samples int32 (H,7), n_rejected)."""
    rng = CvRNG(seed)
    n = len(m1)
    out = []
    rejected = 0
    for _ in range(n_samples):
        found = None
        for _att in range(max_attempts):
            idx = draw_subset(rng, n, 7)
            if fmat_check_subset(m1, m2, idx):
                found = idx
                break
            rejected += 1
        if found is None:
            break
        out.append(found)
    return np.asarray(out, np.int32).reshape(-1, 7), rejected


def fmat_7point_models(s1, s2):
    """OpenCV's 7-point solver on one minimal sample -> list of 3x3 F (1..3)."""
    F, _ = cv2.findFundamentalMat(s1.astype(np.float32), s2.astype(np.float32), cv2.FM_7POINT)
    if F is None:
        return []
    F = np.asarray(F, np.float64).reshape(-1, 3, 3)
    return [F[i] for i in range(F.shape[0])]


def fmat_error(F, m1, m2):
    """FMEstimatorCallback::computeError: max of the two squared point-to-epipolar
    -line distances, double math, stored float32."""
    F = np.asarray(F, np.float64)
    x1 = m1[:, 0].astype(np.float64)
    y1 = m1[:, 1].astype(np.float64)
    x2 = m2[:, 0].astype(np.float64)
    y2 = m2[:, 1].astype(np.float64)
    a = F[0, 0] * x1 + F[0, 1] * y1 + F[0, 2]
    b = F[1, 0] * x1 + F[1, 1] * y1 + F[1, 2]
    c = F[2, 0] * x1 + F[2, 1] * y1 + F[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        s2 = 1.0 / (a * a + b * b)
        d2 = x2 * a + y2 * b + c
        a = F[0, 0] * x2 + F[1, 0] * y2 + F[2, 0]
        b = F[0, 1] * x2 + F[1, 1] * y2 + F[2, 1]
        c = F[0, 2] * x2 + F[1, 2] * y2 + F[2, 2]
        s1 = 1.0 / (a * a + b * b)
        d1 = x1 * a + y1 * b + c
        err = np.maximum(d1 * d1 * s1, d2 * d2 * s2)
    return err.astype(np.float32)


def fmat_lmeds(m1, m2, conf=0.99, max_iters=1000):
    """findFundamentalMat(m1, m2, FM_RANSAC, ...) for 8 <= N <= 14 points: OpenCV does not run RANSAC there but
    LMeDSPointSetRegistrator::run (calib3d fundam.cpp / ptsetreg.cpp): a fixed number of 7-point samples
    (RANSACUpdateNumIters(conf, 0.45, 7, maxIters)), the model with the smallest MEDIAN error wins (first one on
    ties), sigma = 2.5*1.4826*(1 + 5/(N-7))*sqrt(median) (>= 0.001), mask = err <= sigma^2.
    For N <= 13 the median is taken among the seven sample points themselves (errors ~1e-27): the winner is decided
    by rounding noise, in OpenCV too.  Returns (F or None, mask)."""
    m1 = np.ascontiguousarray(m1, np.float32)
    m2 = np.ascontiguousarray(m2, np.float32)
    n = len(m1)
    assert 8 <= n <= 14
    rng = CvRNG(RANSAC_SEED)
    niters = ransac_update_num_iters(conf, 0.45, 7, max_iters)
    best, min_med = None, np.inf
    for _ in range(niters):
        idx = None
        for _att in range(10000):
            cand = draw_subset(rng, n, 7)
            if fmat_check_subset(m1, m2, cand):
                idx = cand
                break
        if idx is None:
            break
        for F in fmat_7point_models(m1[idx], m2[idx]):
            s_err = np.sort(fmat_error(F, m1, m2))
            med = float(s_err[n // 2]) if n % 2 else float(np.float32(s_err[n // 2 - 1] + s_err[n // 2])) * 0.5
            if med < min_med:
                min_med, best = med, F
    if best is None:
        return None, np.zeros(n, np.uint8)
    sigma = max(2.5 * 1.4826 * (1 + 5.0 / (n - 7)) * np.sqrt(min_med), 0.001)
    mask = (fmat_error(best, m1, m2) <= np.float32(sigma * sigma)).astype(np.uint8)
    if mask.sum() < 7:
        return None, np.zeros(n, np.uint8)
    return best, mask


def fmat_ransac(m1, m2, thr, conf, max_iters=1000, samples=None, exhaustive=False):
    """findFundamentalMat(m1, m2, FM_RANSAC, thr, conf) for N >= 15 points.

    samples: optional (H,7) int32 replay list (accepted subsets, in order);
    when given, the RNG is not used.  Returns dict(F, mask, n_iters, samples,
    counts [(sample, model, good)], best=(sample, model)).
    exhaustive: also score the samples past the adaptive stop (for per-hypothesis
    diffs); the returned F/mask keep the early-exit semantics.
    """
    m1 = np.ascontiguousarray(m1, np.float32)
    m2 = np.ascontiguousarray(m2, np.float32)
    n = len(m1)
    assert n >= 8
    t = np.float32(thr * thr)
    rng = CvRNG(RANSAC_SEED)
    niters = max(max_iters, 1)
    best_good = 0
    best_mask = np.zeros(n, np.uint8)
    best_F = None
    best = (-1, -1)
    used = []
    counts = []
    it = 0
    limit = niters
    while it < (limit if exhaustive else niters):
        if samples is not None:
            if it >= len(samples):
                break
            idx = list(samples[it])
        else:
            idx = None
            for _att in range(10000):
                cand = draw_subset(rng, n, 7)
                if fmat_check_subset(m1, m2, cand):
                    idx = cand
                    break
            if idx is None:
                break
        used.append(idx)
        models = fmat_7point_models(m1[idx], m2[idx])
        for mi, F in enumerate(models):
            err = fmat_error(F, m1, m2)
            mask = (err <= t)
            good = int(mask.sum())
            counts.append((it, mi, good))
            if it < niters and good > max(best_good, 6):
                best_good = good
                best_mask = mask.astype(np.uint8)
                best_F = F.copy()
                best = (it, mi)
                niters = ransac_update_num_iters(conf, (n - good) / n, 7, niters)
        it += 1
    return dict(F=best_F, mask=best_mask, n_iters=niters, samples=np.asarray(used, np.int32).reshape(-1, 7),
                counts=counts, best=best, good=best_good)


# ----------------------------------------------------------------------------
# PnP
# ----------------------------------------------------------------------------
def pnp_error(rvec, tvec, K, p3d, p2d):
    """PnPRansacCallback::computeError: projectPoints (double) -> float32 ->
    squared distance accumulated in float32."""
    proj, _ = cv2.projectPoints(p3d.reshape(-1, 1, 3).astype(np.float32), rvec, tvec, K, np.zeros((4, 1)))
    proj = proj.reshape(-1, 2).astype(np.float32)
    d = (p2d.astype(np.float32) - proj).astype(np.float32)
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)


def pnp_epnp_minimal(p3d5, p2d5, K):
    """One RANSAC hypothesis: cv2.solvePnP(5 points, SOLVEPNP_EPNP)."""
    ok, rvec, tvec = cv2.solvePnP(p3d5.reshape(-1, 1, 3).astype(np.float32),
                                  p2d5.reshape(-1, 1, 2).astype(np.float32), K, np.zeros((4, 1)),
                                  flags=cv2.SOLVEPNP_EPNP)
    return ok, rvec, tvec


def pnp_ransac(p3d, p2d, K, iters, thr, conf, samples=None, exhaustive=False, refine=True):
    """solvePnPRansac(p3d, p2d, K, 0, rvec, tvec, false, iters, thr, conf, inliers)
    for N > 5 points (EPnP-5 minimal solver, SOLVEPNP_ITERATIVE refinement).

    Returns dict(ok, rvec, tvec, inliers, best_rvec, best_tvec, n_iters, samples,
    counts [good per sample], hyp [(rvec,tvec) per sample], best).
    """
    p3d = np.ascontiguousarray(p3d, np.float32)
    p2d = np.ascontiguousarray(p2d, np.float32)
    n = len(p3d)
    assert n > 5
    t = np.float32(thr * thr)
    rng = CvRNG(RANSAC_SEED)
    niters = max(iters, 1)
    limit = niters
    best_good = 0
    best_mask = None
    best_model = None
    best = -1
    used = []
    counts = []
    hyp = []
    it = 0
    while it < (limit if exhaustive else niters):
        if samples is not None:
            if it >= len(samples):
                break
            idx = list(samples[it])
        else:
            idx = draw_subset(rng, n, 5)
        used.append(idx)
        ok, rvec, tvec = pnp_epnp_minimal(p3d[idx], p2d[idx], K)
        if not ok:
            counts.append(-1)
            hyp.append(None)
            it += 1
            continue
        err = pnp_error(rvec, tvec, K, p3d, p2d)
        mask = err <= t
        good = int(mask.sum())
        counts.append(good)
        hyp.append((rvec.ravel().copy(), tvec.ravel().copy()))
        if it < niters and good > max(best_good, 4):
            best_good = good
            best_mask = mask.copy()
            best_model = (rvec.copy(), tvec.copy())
            best = it
            niters = ransac_update_num_iters(conf, (n - good) / n, 5, niters)
        it += 1
    out = dict(ok=best_model is not None, rvec=None, tvec=None, inliers=np.zeros(0, np.int32),
               best_rvec=None, best_tvec=None, n_iters=niters,
               samples=np.asarray(used, np.int32).reshape(-1, 5), counts=np.asarray(counts, np.int32),
               hyp=hyp, best=best, good=best_good)
    if best_model is None:
        return out
    inl = np.nonzero(best_mask)[0].astype(np.int32)
    out["inliers"] = inl
    out["best_rvec"] = best_model[0].ravel().copy()
    out["best_tvec"] = best_model[1].ravel().copy()
    if refine:
        o = p3d[inl].astype(np.float64).reshape(-1, 1, 3)
        i2 = p2d[inl].astype(np.float64).reshape(-1, 1, 2)
        ok, rvec, tvec = cv2.solvePnP(o, i2, K, np.zeros((4, 1)), best_model[0].copy(), best_model[1].copy(),
                                      True, cv2.SOLVEPNP_ITERATIVE)
        out["rvec"] = rvec.ravel().copy()
        out["tvec"] = tvec.ravel().copy()
    return out
