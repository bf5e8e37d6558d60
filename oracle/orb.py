"""ORB::detectAndCompute as the reference's loop detector calls it (reference src/optimizationStuff.cpp:49-56),
restated stage by stage.  TEST INFRASTRUCTURE ONLY.

OpenCV is a third-party dependency the reference does not vendor; the parity pin is cv2 4.13.0 as importable in this
image.  Every stage below is pinned against what cv2 exposes of it (tests/test_oracle_orb.py): the
INTER_LINEAR_EXACT pyramid (cv2.resize), FAST-9/16 with suppression (cv2.FastFeatureDetector: positions, order,
scores), the Harris ranking response and the IC_Angle orientation (KeyPoint::response / angle of cv2.ORB.detect),
the smoothing ORB applies before sampling and the rBRIEF descriptors (cv2.ORB.compute on caller-made keypoints), and
the assembled detect_and_compute (cv2.ORB.detectAndCompute: per-octave keypoint sets and descriptors, bit for bit).

Three things had to be recovered from cv2 itself, because OpenCV's source is not in the reference tree:

* PATTERN, the 256 rBRIEF test pairs (orb.cpp: bit_pattern_31_).  cv2 does not export the table.
  tests/golden/recover_orb_pattern.py recovers it: a keypoint at angle 0 on a constant image with ONE bright (dark)
  pixel at offset p sets exactly the bits whose second (first) test point lies within the 7x7 smoothing footprint of
  p; fitting the pair positions to the observed bit maps of all 37 x 37 offsets under the measured impulse response
  gives every pair uniquely and with zero residual.  tests/test_oracle_orb.py re-checks the table against live cv2
  with the forward model.
* the smoothing: ORB blurs each pyramid level in place with GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101); the level
  is a sub-matrix of the pyramid buffer, for which OpenCV does NOT take its bit-exact fixed-point Gaussian but the
  generic float separable filter (`cv2.sepFilter2D` with `getGaussianKernel(7, 2, CV_32F)` reproduces ORB's bits,
  `cv2.GaussianBlur` on the whole image does not: 2,862 of 466,616 pixels differ by 1 on the golden frame).
* the evaluation order of that float filter in this build (AVX2/FMA dispatch): rows  s = k0*p0, s = fma(k_j, p_j, s)
  for j = 1..6;  columns  t = k3*s0, t = fma(k_{3+j}, s_{+j} + s_{-j}, t) for j = 1..3;  result = rint(t).  Of the
  eight fused/unfused/symmetric variants this is the one with zero differing pixels (the others differ at 1-3 pixels
  per frame), so a cv2 built for another instruction set may differ from this restatement in isolated bits.
"""
import numpy as np

# (x1, y1, x2, y2) of test k: bit k of the descriptor = I(p1) < I(p2)
PATTERN = np.array([
    (8, -3, 9, 5), (4, 2, 7, -12), (-11, 9, -8, 2), (7, -12, 12, -13),
    (2, -13, 2, 12), (1, -7, 1, 6), (-2, -10, -2, -4), (-13, -13, -11, -8),
    (-13, -3, -12, -9), (10, 4, 11, 9), (-13, -8, -8, -9), (-11, 7, -9, 12),
    (7, 7, 12, 6), (-4, -5, -3, 0), (-13, 2, -12, -3), (-9, 0, -7, 5),
    (12, -6, 12, -1), (-3, 6, -2, 12), (-6, -13, -4, -8), (11, -13, 12, -8),
    (4, 7, 5, 1), (5, -3, 10, -3), (3, -7, 6, 12), (-8, -7, -6, -2),
    (-2, 11, -1, -10), (-13, 12, -8, 10), (-7, 3, -5, -3), (-4, 2, -3, 7),
    (-10, -12, -6, 11), (5, -12, 6, -7), (5, -6, 7, -1), (1, 0, 4, -5),
    (9, 11, 11, -13), (4, 7, 4, 12), (2, -1, 4, 4), (-4, -12, -2, 7),
    (-8, -5, -7, -10), (4, 11, 9, 12), (0, -8, 1, -13), (-13, -2, -8, 2),
    (-3, -2, -2, 3), (-6, 9, -4, -9), (8, 12, 10, 7), (0, 9, 1, 3),
    (7, -5, 11, -10), (-13, -6, -11, 0), (10, 7, 12, 1), (-6, -3, -6, 12),
    (10, -9, 12, -4), (-13, 8, -8, -12), (-13, 0, -8, -4), (3, 3, 7, 8),
    (5, 7, 10, -7), (-1, 7, 1, -12), (3, -10, 5, 6), (2, -4, 3, -10),
    (-13, 0, -13, 5), (-13, -7, -12, 12), (-13, 3, -11, 8), (-7, 12, -4, 7),
    (6, -10, 12, 8), (-9, -1, -7, -6), (-2, -5, 0, 12), (-12, 5, -7, 5),
    (3, -10, 8, -13), (-7, -7, -4, 5), (-3, -2, -1, -7), (2, 9, 5, -11),
    (-11, -13, -5, -13), (-1, 6, 0, -1), (5, -3, 5, 2), (-4, -13, -4, 12),
    (-9, -6, -9, 6), (-12, -10, -8, -4), (10, 2, 12, -3), (7, 12, 12, 12),
    (-7, -13, -6, 5), (-4, 9, -3, 4), (7, -1, 12, 2), (-7, 6, -5, 1),
    (-13, 11, -12, 5), (-3, 7, -2, -6), (7, -8, 12, -7), (-13, -7, -11, -12),
    (1, -3, 12, 12), (2, -6, 3, 0), (-4, 3, -2, -13), (-1, -13, 1, 9),
    (7, 1, 8, -6), (1, -1, 3, 12), (9, 1, 12, 6), (-1, -9, -1, 3),
    (-13, -13, -10, 5), (7, 7, 10, 12), (12, -5, 12, 9), (6, 3, 7, 11),
    (5, -13, 6, 10), (2, -12, 2, 3), (3, 8, 4, -6), (2, 6, 12, -13),
    (9, -12, 10, 3), (-8, 4, -7, 9), (-11, 12, -4, -6), (1, 12, 2, -8),
    (6, -9, 7, -4), (2, 3, 3, -2), (6, 3, 11, 0), (3, -3, 8, -8),
    (7, 8, 9, 3), (-11, -5, -6, -4), (-10, 11, -5, 10), (-5, -8, -3, 12),
    (-10, 5, -9, 0), (8, -1, 12, -6), (4, -6, 6, -11), (-10, 12, -8, 7),
    (4, -2, 6, 7), (-2, 0, -2, 12), (-5, -8, -5, 2), (7, -6, 10, 12),
    (-9, -13, -8, -8), (-5, -13, -5, -2), (8, -8, 9, -13), (-9, -11, -9, 0),
    (1, -8, 1, -2), (7, -4, 9, 1), (-2, 1, -1, -4), (11, -6, 12, -11),
    (-12, -9, -6, 4), (3, 7, 7, 12), (5, 5, 10, 8), (0, -4, 2, 8),
    (-9, 12, -5, -13), (0, 7, 2, 12), (-1, 2, 1, 7), (5, 11, 7, -9),
    (3, 5, 6, -8), (-13, -4, -8, 9), (-5, 9, -3, -3), (-4, -7, -3, -12),
    (6, 5, 8, 0), (-7, 6, -6, 12), (-13, 6, -5, -2), (1, -10, 3, 10),
    (4, 1, 8, -4), (-2, -2, 2, -13), (2, -12, 12, 12), (-2, -13, 0, -6),
    (4, 1, 9, 3), (-6, -10, -3, -5), (-3, -13, -1, 1), (7, 5, 12, -11),
    (4, -2, 5, -7), (-13, 9, -9, -5), (7, 1, 8, 6), (7, -8, 7, 6),
    (-7, -4, -7, 1), (-8, 11, -7, -8), (-13, 6, -12, -8), (2, 4, 3, 9),
    (10, -5, 12, 3), (-6, -5, -6, 7), (8, -3, 9, -8), (2, -12, 2, 8),
    (-11, -2, -10, 3), (-12, -13, -7, -9), (-11, 0, -10, -5), (5, -3, 11, 8),
    (-2, -13, -1, 12), (-1, -8, 0, 9), (-13, -11, -12, -5), (-10, -2, -10, 11),
    (-3, 9, -2, -13), (2, -3, 3, 2), (-9, -13, -4, 0), (-4, 6, -3, -10),
    (-4, 12, -2, -7), (-6, -11, -4, 9), (6, -3, 6, 11), (-13, 11, -5, 5),
    (11, 11, 12, 6), (7, -5, 12, -2), (-1, 12, 0, 7), (-4, -8, -3, -2),
    (-7, 1, -6, 7), (-13, -12, -8, -13), (-7, -2, -6, -8), (-8, 5, -6, -9),
    (-5, -1, -4, 5), (-13, 7, -8, 10), (1, 5, 5, -13), (1, 0, 10, -13),
    (9, 12, 10, -1), (5, -8, 10, -9), (-1, 11, 1, -13), (-9, -3, -6, 2),
    (-1, -10, 1, 12), (-13, 1, -8, -10), (8, -11, 10, -6), (2, -13, 3, -6),
    (7, -13, 12, -9), (-10, -10, -5, -7), (-10, -8, -8, -13), (4, -6, 8, 5),
    (3, 12, 8, -13), (-4, 2, -3, -3), (5, -13, 10, -12), (4, -13, 5, -1),
    (-9, 9, -4, 3), (0, 3, 3, -9), (-12, 1, -6, 1), (3, 2, 4, -8),
    (-10, -10, -10, 9), (8, -13, 12, 12), (-8, -12, -6, -5), (2, 2, 3, 7),
    (10, 6, 11, -8), (6, 8, 8, -12), (-7, 10, -6, 5), (-3, -9, -3, 9),
    (-1, -13, -1, 5), (-3, -7, -3, 4), (-8, -2, -8, 3), (4, 2, 12, 12),
    (2, -5, 3, 11), (6, -9, 11, -13), (3, -1, 7, 12), (11, -1, 12, 4),
    (-3, 0, -3, 6), (4, -11, 4, 12), (2, -4, 2, 1), (-10, -6, -8, 1),
    (-13, 7, -11, 1), (-13, 12, -11, -13), (6, 0, 11, -13), (0, -1, 1, 4),
    (-13, 3, -9, -2), (-9, 8, -6, -3), (-13, -6, -8, -2), (5, -9, 8, 10),
    (2, 7, 3, -9), (-1, -6, -1, -1), (9, 5, 11, -2), (11, -3, 12, -8),
    (3, 0, 3, 5), (-1, 4, 0, 10), (3, -6, 4, 5), (-13, 0, -10, 5),
    (5, 8, 12, 11), (8, 9, 9, -6), (7, -4, 8, -12), (-10, 4, -10, 9),
    (7, 3, 12, 4), (9, -7, 10, -2), (7, 0, 12, -2), (-1, -6, 0, -11),
], np.int32)

HALF_PATCH_REACH = 19      # no rotated test point leaves [-19, 19]^2 (|p| <= 13 * sqrt(2))


def gaussian_kernel_7_2():
    """getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) normalised, computed in double, stored as float."""
    x = np.arange(-3, 4, dtype=np.float64)
    k = np.exp(-(x * x) / 8.0)
    return (k / k.sum()).astype(np.float32)


def _fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def smooth(img):
    """The image ORB samples its descriptors from (see the module docstring for the evaluation order)."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    k = gaussian_kernel_7_2()
    p = np.pad(img, 3, mode="reflect").astype(np.float32)
    s = (p[:, 0:w] * k[0]).astype(np.float32)
    for j in range(1, 7):
        s = _fma(p[:, j:j + w], k[j], s)
    t = (s[3:3 + h] * k[3]).astype(np.float32)
    for j in range(1, 4):
        t = _fma((s[3 + j:3 + j + h] + s[3 - j:3 - j + h]).astype(np.float32), k[3 + j], t)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def smooth_call_through(img):
    import cv2
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    return cv2.sepFilter2D(np.ascontiguousarray(img), -1, k, k, borderType=cv2.BORDER_REFLECT_101)


def describe(img, xy, angle_deg, smoothed=None):
    """rBRIEF descriptors (n x 32 uint8) of keypoints xy (n x 2 float32) with angles in degrees on ONE level.
    orb.cpp computeOrbDescriptors: angle *= (float)(CV_PI / 180); a = (float)cos(angle), b = (float)sin(angle);
    test point (x, y) -> (cvRound(x*a - y*b), cvRound(x*b + y*a)) around (cvRound(pt.x), cvRound(pt.y))."""
    sm = smooth(img) if smoothed is None else smoothed
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    ang = np.asarray(angle_deg, np.float32).reshape(-1) * np.float32(np.pi / 180.0)
    a = np.cos(ang.astype(np.float64)).astype(np.float32)[:, None, None]
    b = np.sin(ang.astype(np.float64)).astype(np.float32)[:, None, None]
    px = PATTERN[:, [0, 2]].astype(np.float32)[None]
    py = PATTERN[:, [1, 3]].astype(np.float32)[None]
    ix = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
    iy = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
    cx = np.rint(xy[:, 0]).astype(np.int64)[:, None, None]
    cy = np.rint(xy[:, 1]).astype(np.int64)[:, None, None]
    v = sm[cy + iy, cx + ix].astype(np.int32)
    return np.packbits((v[..., 0] < v[..., 1]).astype(np.uint8), axis=1, bitorder="little")


def describe_call_through(img, xy, angle_deg):
    """cv2.ORB_create().compute on caller-made octave-0 keypoints (size 31); cv2 drops keypoints closer than 31 px
    to the border, so the caller keeps them inside."""
    import cv2
    kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, 0, -1)
           for (x, y), a in zip(np.asarray(xy, np.float32), np.asarray(angle_deg, np.float32))]
    out_kp, desc = cv2.ORB_create().compute(np.ascontiguousarray(img), kps)
    assert len(out_kp) == len(kps), "cv2 dropped keypoints (too close to the border)"
    return desc


# ---------------------------------------------------------------------------------------------- orientation
def umax_table(half_patch=15):
    """orb.cpp: the half-width of every row of the circular patch"""
    import math
    s2 = float(np.float32(np.sqrt(np.float32(2.0))))
    vmax = int(math.floor(half_patch * s2 / 2 + 1))
    vmin = int(math.ceil(half_patch * s2 / 2))
    u = [0] * (half_patch + 2)
    for v in range(vmax + 1):
        u[v] = int(np.rint(math.sqrt(half_patch * half_patch - v * v)))
    v0 = 0
    for v in range(half_patch, vmin - 1, -1):
        while u[v0] == u[v0 + 1]:
            v0 += 1
        u[v] = v0
        v0 += 1
    return u[:half_patch + 1]


def fast_atan2(y, x):
    """cv::fastAtan2 (degrees), scalar float path: odd 7th-order polynomial in min/max, no contraction."""
    f = np.float32
    y = np.asarray(y, f)
    x = np.asarray(x, f)
    s = f(180.0 / np.pi)
    p1, p3 = f(0.9997878412794807) * s, f(-0.3258083974640975) * s
    p5, p7 = f(0.1555786518463281) * s, f(-0.04432655554792128) * s
    ax, ay = np.abs(x), np.abs(y)
    eps = f(2.220446049250313e-16)
    with np.errstate(all="ignore"):
        c = np.where(ax >= ay, ay / (ax + eps), ax / (ay + eps)).astype(f)
    c2 = (c * c).astype(f)
    a = ((((((p7 * c2).astype(f) + p5).astype(f) * c2).astype(f) + p3).astype(f) * c2).astype(f) + p1).astype(f)
    a = (a * c).astype(f)
    a = np.where(ax >= ay, a, (f(90.0) - a).astype(f))
    a = np.where(x < 0, (f(180.0) - a).astype(f), a)
    a = np.where(y < 0, (f(360.0) - a).astype(f), a)
    return a.astype(f)


def ic_angle(img, xy, half_patch=15):
    """orb.cpp ICAngles on the UNSMOOTHED level: m_10 = sum u*I, m_01 = sum v*I over the circular patch around the
    rounded keypoint position, angle = fastAtan2((float)m_01, (float)m_10) in degrees."""
    I = np.asarray(img, np.uint8).astype(np.int64)
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    cx = np.rint(xy[:, 0]).astype(np.int64)
    cy = np.rint(xy[:, 1]).astype(np.int64)
    um = umax_table(half_patch)
    m10 = np.zeros(len(xy), np.int64)
    m01 = np.zeros(len(xy), np.int64)
    for v in range(-half_patch, half_patch + 1):
        d = um[abs(v)]
        for u in range(-d, d + 1):
            val = I[cy + v, cx + u]
            m10 += u * val
            m01 += v * val
    return fast_atan2(m01.astype(np.float32), m10.astype(np.float32))


def detect_call_through(img, nfeatures=30000, fast_threshold=3, octave=0):
    """cv2.ORB.detect: (xy, angle) of the keypoints of one octave -- the pin for ic_angle"""
    import cv2
    kps = [k for k in cv2.ORB_create(nfeatures=nfeatures, fastThreshold=fast_threshold).detect(np.ascontiguousarray(img), None)
           if k.octave == octave]
    return (np.array([k.pt for k in kps], np.float32).reshape(-1, 2), np.array([k.angle for k in kps], np.float32))


# ---------------------------------------------------------------------------------------------- ranking
def harris_response(img, xy, block=7, k=0.04):
    """orb.cpp HarrisResponses: 3x3 Sobel-like gradients summed over a block x block window around the rounded
    keypoint (integers), response = (a*b - c*c - k*(a+b)^2) * scale^4 in float, scale = 1 / (4 * block * 255).
    This is the `response` cv2.ORB.detect reports and ranks by."""
    I = np.asarray(img, np.uint8).astype(np.int64)
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    r = block // 2
    x0 = np.rint(xy[:, 0]).astype(np.int64)
    y0 = np.rint(xy[:, 1]).astype(np.int64)
    a = np.zeros(len(xy), np.int64)
    b = np.zeros(len(xy), np.int64)
    c = np.zeros(len(xy), np.int64)
    for i in range(block):
        for j in range(block):
            y, x = y0 - r + i, x0 - r + j
            ix = (I[y, x + 1] - I[y, x - 1]) * 2 + (I[y - 1, x + 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y + 1, x - 1])
            iy = (I[y + 1, x] - I[y - 1, x]) * 2 + (I[y + 1, x - 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y - 1, x + 1])
            a += ix * ix
            b += iy * iy
            c += ix * iy
    f = np.float32
    scale = f(1.0) / f(f(4 * block) * f(255.0))
    s4 = f(f(f(scale * scale) * scale) * scale)
    af, bf, cf = a.astype(f), b.astype(f), c.astype(f)
    t = ((af * bf).astype(f) - (cf * cf).astype(f)).astype(f)
    apb = (af + bf).astype(f)
    t = (t - ((f(k) * apb).astype(f) * apb).astype(f)).astype(f)
    return (t * s4).astype(f)


def detect_call_through_full(img, nfeatures=30000, fast_threshold=3, octave=0):
    """(xy, angle, response) of cv2.ORB.detect's keypoints of one octave"""
    import cv2
    kps = [k for k in cv2.ORB_create(nfeatures=nfeatures, fastThreshold=fast_threshold).detect(np.ascontiguousarray(img), None)
           if k.octave == octave]
    return (np.array([k.pt for k in kps], np.float32).reshape(-1, 2), np.array([k.angle for k in kps], np.float32),
            np.array([k.response for k in kps], np.float32))


# ---------------------------------------------------------------------------------------------- detection
FAST_CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
               (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def fast9(img, threshold=20, nonmax=True):
    """cv::FAST (TYPE_9_16), the detector ORB runs on every pyramid level (orb.cpp computeKeyPoints:
    FastFeatureDetector::create(fastThreshold, true)).  A pixel is a corner when 9 contiguous pixels of the 16-pixel
    circle are all darker than v - t or all brighter than v + t; its score (cornerScore<16>) is the largest such
    margin minus 1; non-maximum suppression keeps a corner whose score is strictly greater than its 8 neighbours'
    (non-corners count as 0).  Returns (xy float32 in raster order, score float32, score map)."""
    I = np.asarray(img, np.uint8).astype(np.int32)
    h, w = I.shape
    v = I[3:h - 3, 3:w - 3]
    d = np.stack([v - I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in FAST_CIRCLE], 0)
    d2 = np.concatenate([d, d[:8]], 0)
    dark = np.stack([d2[k:k + 9].min(0) for k in range(16)], 0).max(0)
    bright = np.stack([(-d2[k:k + 9]).min(0) for k in range(16)], 0).max(0)
    best = np.maximum(dark, bright)
    S = np.zeros((h, w), np.int32)
    S[3:h - 3, 3:w - 3] = np.where(best > threshold, best - 1, 0)
    if nonmax:
        c = S[1:-1, 1:-1]
        keep = c > 0
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx or dy:
                    keep &= c > S[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
        K = np.zeros((h, w), bool)
        K[1:-1, 1:-1] = keep
    else:
        K = S > 0
    ys, xs = np.nonzero(K)
    return np.c_[xs, ys].astype(np.float32), S[ys, xs].astype(np.float32), S


def fast9_call_through(img, threshold=20, nonmax=True):
    import cv2
    det = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=nonmax,
                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(np.ascontiguousarray(img), None)
    return (np.array([k.pt for k in kps], np.float32).reshape(-1, 2), np.array([k.response for k in kps], np.float32))


# ---------------------------------------------------------------------------------------------- pyramid + glue
def _exact_coeffs(srcsize, dstsize):
    """resize.cpp interpolationLinear::getCoeffs for 8-bit images: source offset and the 8.8 fixed-point weight of
    its right/lower neighbour; positions that fall off the source take the border pixel (weight 0)."""
    import math
    inv = np.float64(dstsize) / np.float64(srcsize)
    scale = np.float64(1.0) / inv
    ofs = np.zeros(dstsize, np.int64)
    c1 = np.zeros(dstsize, np.int64)
    for v in range(dstsize):
        fval = scale * (np.float64(v) + 0.5) - 0.5
        ival = math.floor(fval)
        if ival >= 0 and srcsize > 1:
            if ival < srcsize - 1:
                ofs[v] = ival
                c1[v] = int(np.rint((fval - ival) * 256.0))
            else:
                ofs[v] = srcsize - 1
        # else: before the first pixel -> offset 0, weight 0
    return ofs, c1


def resize_linear_exact(src, dw, dh):
    """cv::resize(..., INTER_LINEAR_EXACT) for 8-bit single-channel images: horizontal pass in 8.8 fixed point,
    vertical pass in 16.16, rounded half up."""
    src = np.asarray(src, np.uint8)
    sh, sw = src.shape
    s = src.astype(np.int64)
    ox, cx = _exact_coeffs(sw, dw)
    oy, cy = _exact_coeffs(sh, dh)
    hrow = s[:, ox] * (256 - cx) + s[:, np.minimum(ox + 1, sw - 1)] * cx
    out = hrow[oy] * (256 - cy)[:, None] + hrow[np.minimum(oy + 1, sh - 1)] * cy[:, None]
    return ((out + 32768) >> 16).astype(np.uint8)


SCALE_FACTOR = float(np.float32(1.2))      # ORB::create takes a float 1.2f and keeps it in a double member


def level_scales(nlevels=8, scale_factor=SCALE_FACTOR):
    return [np.float32(np.float64(scale_factor) ** lvl) for lvl in range(nlevels)]     # (float)std::pow(scaleFactor, level)


def build_pyramid(img, nlevels=8, scale_factor=SCALE_FACTOR):
    """orb.cpp detectAndCompute: level k = resize(level k-1, cvRound(size / scale_k), INTER_LINEAR_EXACT)"""
    img = np.asarray(img, np.uint8)
    out = [img]
    for sc in level_scales(nlevels, scale_factor)[1:]:
        dw = int(np.rint(np.float32(img.shape[1]) / sc))
        dh = int(np.rint(np.float32(img.shape[0]) / sc))
        out.append(resize_linear_exact(out[-1], dw, dh))
    return out


def features_per_level(nfeatures=500, nlevels=8, scale_factor=SCALE_FACTOR):
    f = np.float32
    factor = f(1.0 / scale_factor)
    nd = f(f(nfeatures) * f(f(1) - factor)) / f(f(1) - f(np.float64(factor) ** np.float64(nlevels)))
    nd = f(nd)
    out, tot = [], 0
    for _ in range(nlevels - 1):
        n = int(np.rint(nd))
        out.append(n)
        tot += n
        nd = f(nd * factor)
    out.append(max(nfeatures - tot, 0))
    return out


def retain_best(resp, n):
    """KeyPointsFilter::retainBest: indices of all keypoints whose response is >= the n-th largest (ties stay)"""
    resp = np.asarray(resp)
    if n <= 0:
        return np.zeros(0, np.int64)
    if len(resp) <= n:
        return np.arange(len(resp))
    thr = np.sort(resp)[::-1][n - 1]
    return np.nonzero(resp >= thr)[0]


def detect_and_compute(img, nfeatures=500, nlevels=8, scale_factor=SCALE_FACTOR, edge_threshold=31, fast_threshold=20):
    """ORB::detectAndCompute with the reference's defaults (ORB::create(), src/optimizationStuff.cpp:49-56): per level
    FAST -> border filter -> retainBest(2n) by FAST score -> Harris response -> retainBest(n) -> IC_Angle -> smoothing ->
    rBRIEF; positions scaled back to level 0.  Returns a dict of arrays sorted by (octave, y, x): cv2's own order
    inside a level is whatever std::nth_element leaves and is not part of the contract."""
    pyr = build_pyramid(img, nlevels, scale_factor)
    scales = level_scales(nlevels, scale_factor)
    quota = features_per_level(nfeatures, nlevels, scale_factor)
    out = dict(xy=[], octave=[], response=[], angle=[], desc=[], level_xy=[])
    for lvl, (im, sf, n) in enumerate(zip(pyr, scales, quota)):
        h, w = im.shape
        if w <= 2 * edge_threshold or h <= 2 * edge_threshold:      # the border filter leaves nothing
            continue
        xy, sc, _ = fast9(im, fast_threshold, True)
        inside = (xy[:, 0] >= edge_threshold) & (xy[:, 0] < w - edge_threshold) & \
                 (xy[:, 1] >= edge_threshold) & (xy[:, 1] < h - edge_threshold)
        xy, sc = xy[inside], sc[inside]
        keep = retain_best(sc, 2 * n)
        xy = xy[keep]
        resp = harris_response(im, xy)
        keep = retain_best(resp, n)
        xy, resp = xy[keep], resp[keep]
        if len(xy) == 0:
            continue
        ang = ic_angle(im, xy)
        desc = describe(im, xy, ang)
        order = np.lexsort((xy[:, 0], xy[:, 1]))
        out["level_xy"].append(xy[order])
        out["xy"].append((xy[order] * np.float32(sf)).astype(np.float32))
        out["octave"].append(np.full(len(xy), lvl, np.int32))
        out["response"].append(resp[order])
        out["angle"].append(ang[order])
        out["desc"].append(desc[order])
    return {k: (np.concatenate(v) if v else np.zeros((0,))) for k, v in out.items()}


def detect_and_compute_call_through(img, nfeatures=500):
    """cv2.ORB_create(nfeatures).detectAndCompute, sorted like detect_and_compute"""
    import cv2
    kps, desc = cv2.ORB_create(nfeatures=nfeatures).detectAndCompute(np.ascontiguousarray(img), None)
    xy = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    octv = np.array([k.octave for k in kps], np.int32)
    resp = np.array([k.response for k in kps], np.float32)
    ang = np.array([k.angle for k in kps], np.float32)
    order = np.lexsort((xy[:, 0], xy[:, 1], octv))
    return dict(xy=xy[order], octave=octv[order], response=resp[order], angle=ang[order],
                desc=(desc[order] if desc is not None else np.zeros((0, 32), np.uint8)))
