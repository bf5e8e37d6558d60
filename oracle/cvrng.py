"""OpenCV's RNG (multiply-with-carry), restated.  Test infrastructure only.

cv::RNG::next():  state = (uint64)(uint32)state * 4164903690 + (state >> 32);
                  return (uint32)state
cv::RNG::uniform(a, b) for ints:  a + next() % (b - a)

Both RANSAC loops used by the reference (findFundamentalMat at
tracking.cpp:34,75 and solvePnPRansac at keyFrameManagement.cpp:84,88) seed it
with 0xFFFFFFFFFFFFFFFF (cv::RNG rng((uint64)-1) in RANSACPointSetRegistrator).
"""
import numpy as np

CV_RNG_COEFF = 4164903690
RANSAC_SEED = 0xFFFFFFFFFFFFFFFF
_M64 = (1 << 64) - 1


class CvRNG:
    def __init__(self, state=RANSAC_SEED):
        self.state = state & _M64 if state else 0xFFFFFFFF

    def next(self):
        s = self.state
        s = ((s & 0xFFFFFFFF) * CV_RNG_COEFF + (s >> 32)) & _M64
        self.state = s
        return s & 0xFFFFFFFF

    def uniform(self, a, b):
        return a + self.next() % (b - a)


def draw_subset(rng, count, model_points, max_attempts=10000):
    """RANSACPointSetRegistrator::getSubset index drawing: ``model_points``
    distinct indices in draw order, redrawing on duplicates.  Returns a list,
    or None after ``max_attempts`` failed attempts (cannot happen without a
    subset check, kept for shape)."""
    idx = []
    i = 0
    iters = 0
    while i < model_points and iters < max_attempts:
        while True:
            idx_i = rng.uniform(0, count)
            if idx_i not in idx[:i]:
                break
        if len(idx) <= i:
            idx.append(idx_i)
        else:
            idx[i] = idx_i
        i += 1
    return idx if i == model_points else None


def sample_list(count, model_points, n_samples, seed=RANSAC_SEED):
    """The first ``n_samples`` minimal-sample index tuples OpenCV's RANSAC loop
    would draw for a point set of size ``count`` when no subset is rejected."""
    rng = CvRNG(seed)
    out = np.empty((n_samples, model_points), np.int32)
    for h in range(n_samples):
        out[h] = draw_subset(rng, count, model_points)
    return out


def ransac_update_num_iters(p, ep, model_points, max_iters):
    """cv::RANSACUpdateNumIters."""
    import math
    p = max(p, 0.0)
    p = min(p, 1.0)
    ep = max(ep, 0.0)
    ep = min(ep, 1.0)
    num = max(1.0 - p, 2.2250738585072014e-308)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < 2.2250738585072014e-308:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    # cvRound = round half to even
    return int(np.rint(num / denom))
