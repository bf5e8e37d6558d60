"""Recovers OpenCV's 256 rBRIEF test pairs (orb.cpp: bit_pattern_31_) from cv2 itself and prints them in the form
oracle/orb.py embeds.  cv2 does not export the table and OpenCV's source is not in the reference tree.

Method: `cv2.ORB.compute` accepts caller-made keypoints.  On a constant image with ONE bright pixel at offset p from a
keypoint with angle 0, bit k is set iff its second test point sees more of the (7x7-smoothed) impulse than its first;
with ONE dark pixel the roles swap.  The bit maps over all offsets p in [-18, 18]^2 are fitted, per bit, with integer
pair positions under the measured impulse response of ORB's smoothing; every pair is found uniquely with zero residual.

    python tests/golden/recover_orb_pattern.py        (about a minute)
"""
import numpy as np
import cv2

R = 18
C = 50
S = 101


def observe():
    orb = cv2.ORB_create()
    base = np.full((S, S), 128, np.uint8)
    hits = np.zeros((2, 256, 2 * R + 1, 2 * R + 1), bool)     # [dark / bright][bit][dy][dx]
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            for j, val in enumerate((0, 255)):
                im = base.copy()
                im[C + dy, C + dx] = val
                _, d = orb.compute(im, [cv2.KeyPoint(float(C), float(C), 31.0, 0.0, 1.0, 0, -1)])
                hits[j, :, dy + R, dx + R] = np.unpackbits(d[0], bitorder="little").astype(bool)
    return hits


def impulse_response(val):
    """ORB smooths with the generic float separable filter (see oracle/orb.py), i.e. cv2.sepFilter2D."""
    im = np.full((41, 41), 128, np.uint8)
    im[20, 20] = val
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    return cv2.sepFilter2D(im, -1, k, k, borderType=cv2.BORDER_REFLECT_101).astype(np.int32)


def seen(resp, q, ys, xs):
    """smoothed value at test point q when the impulse sits at (xs, ys)"""
    oy, ox = q[1] - ys, q[0] - xs
    inside = (np.abs(oy) <= 20) & (np.abs(ox) <= 20)
    v = np.full(ys.shape, 128, np.int32)
    v[inside] = resp[20 + oy[inside], 20 + ox[inside]]
    return v


def predict(pattern):
    """forward model: the bit maps a table would produce (used by tests/test_oracle_orb.py)"""
    ys, xs = np.mgrid[-R:R + 1, -R:R + 1]
    rd, rb = impulse_response(0), impulse_response(255)
    out = np.zeros((2, 256, 2 * R + 1, 2 * R + 1), bool)
    for k, (x1, y1, x2, y2) in enumerate(pattern):
        out[0, k] = seen(rd, (x1, y1), ys, xs) < seen(rd, (x2, y2), ys, xs)
        out[1, k] = seen(rb, (x1, y1), ys, xs) < seen(rb, (x2, y2), ys, xs)
    return out


def fit(hits):
    ys, xs = np.mgrid[-R:R + 1, -R:R + 1]
    rd, rb = impulse_response(0), impulse_response(255)
    table = np.zeros((256, 4), int)
    for k in range(256):
        c1 = (xs[hits[0, k]].mean(), ys[hits[0, k]].mean())      # dark impulse lights the bit near the FIRST point
        c2 = (xs[hits[1, k]].mean(), ys[hits[1, k]].mean())
        zero = []
        for x1 in range(int(np.floor(c1[0])) - 2, int(np.ceil(c1[0])) + 3):
            for y1 in range(int(np.floor(c1[1])) - 2, int(np.ceil(c1[1])) + 3):
                for x2 in range(int(np.floor(c2[0])) - 2, int(np.ceil(c2[0])) + 3):
                    for y2 in range(int(np.floor(c2[1])) - 2, int(np.ceil(c2[1])) + 3):
                        e = ((seen(rd, (x1, y1), ys, xs) < seen(rd, (x2, y2), ys, xs)) != hits[0, k]).sum() \
                            + ((seen(rb, (x1, y1), ys, xs) < seen(rb, (x2, y2), ys, xs)) != hits[1, k]).sum()
                        if e == 0:
                            zero.append((x1, y1, x2, y2))
        assert len(zero) == 1, (k, zero)
        table[k] = zero[0]
    return table


if __name__ == "__main__":
    t = fit(observe())
    for i in range(0, 256, 4):
        print("    " + " ".join("(%d, %d, %d, %d)," % tuple(t[k]) for k in range(i, i + 4)))
