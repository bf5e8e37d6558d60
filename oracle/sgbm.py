"""Restatement of StereoProcess::stereoMatch / reprojectDisparity (reference src/StereoCV.cpp:21-62,
221-250).  TEST INFRASTRUCTURE ONLY.

The reference converts both frames to gray (`cvtColor(BGR2GRAY)`, StereoCV.cpp:35-36), runs
`StereoSGBM::create(1, 96, 7, 24, 96, 0, 60, 0, 3000, 5)->compute` (StereoCV.cpp:39-53) and reprojects the raw
16x fixed-point disparity with `reprojectImageTo3D` and the Q matrix of `stereoRectify` (StereoCV.cpp:229-232),
keeping points with 0.01 < z <= 5 and flipping y (StereoCV.cpp:237-244).

The arithmetic lives in OpenCV (third party, not vendored, version unpinned by the reference); the parity pin is
cv2 4.13.0 as importable in this image.  `sgbm_call_through` calls cv2 with the reference's arguments;
`sgbm_compute` restates the algorithm of `StereoSGBM` in MODE_SGBM (modules/calib3d/src/stereosgbm.cpp:
calcPixelCostBT, computeDisparitySGBM, then medianBlur 3x3 and filterSpeckles) stage by stage, so that the CUDA
path can be compared per stage, and is PINNED BIT-IDENTICAL to cv2 by tests/test_oracle_sgbm.py.

Stages (all integer):
  1  x-Sobel prefilter clipped to [-ftzero, ftzero] + ftzero (ftzero = max(preFilterCap, 15) | 1); columns 0 and
     width-1 of both the filtered and the raw plane read as ftzero (OpenCV sets the row ends to tab[0]).
  2  Birchfield-Tomasi cost on both planes, the raw one scaled by 1/4:
     pix(y, x, d) for x in [maxD, width), d in [minD, maxD).
  3  box sum over the block, rows replicated at the image border, columns replicated at the border of the
     *computed* x range -> C(y, x, d), int16.
  4  five path costs L_r(p, d) = C + min(L(p-r, d), L(p-r, d-1) + P1, L(p-r, d+1) + P1, min_k L(p-r, k) + P2)
     - min_k L(p-r, k) with L = 0 outside the range: left-to-right, three top-down directions, right-to-left;
     S = sat16(sat16(L0 + L1 + L2 + L3) + L4).  (Whether P2 is subtracted as well only shifts S by 5 * P2; the
     uniqueness test is the one place that sees the shift, and cv2 4.13.0 behaves as written here.)
  5  winner-take-all (first minimum), uniqueness test, right-view disparity by first-come-lowest-cost in
     descending x, parabola sub-pixel step with C integer division, left-right check (disp12MaxDiff <= 0 -> 1).
  6  medianBlur 3x3 (replicated border), filterSpeckles(newVal = (minD - 1) * 16, speckleWindow, 16 * range):
     4-connected components over |difference| <= maxDiff edges, components of <= speckleWindow pixels are cleared.
"""
import numpy as np

DISP_SHIFT = 4
DISP_SCALE = 16
MAX_COST = 32767

# the reference's arguments, StereoCV.cpp:39-50
REF_PARAMS = dict(min_disp=1, num_disp=96, block=7, P1=24, P2=96, disp12_max_diff=0, pre_filter_cap=60,
                  uniqueness=0, speckle_window=3000, speckle_range=5)


def sgbm_call_through(L, R, **kw):
    """cv2.StereoSGBM with the reference's arguments (or overrides): int16 disparity, 16x fixed point."""
    import cv2
    p = dict(REF_PARAMS)
    p.update(kw)
    m = cv2.StereoSGBM_create(p["min_disp"], p["num_disp"], p["block"], p["P1"], p["P2"], p["disp12_max_diff"],
                              p["pre_filter_cap"], p["uniqueness"], p["speckle_window"], p["speckle_range"])
    return m.compute(np.ascontiguousarray(L), np.ascontiguousarray(R))


def prefilter(img, ftzero):
    """Stage 1: (filtered, raw) planes as calcPixelCostBT builds them for every row."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    a = img.astype(np.int32)
    up = np.vstack([a[:1], a[:-1]])
    dn = np.vstack([a[1:], a[-1:]])
    f = np.full((h, w), ftzero, np.int32)
    raw = np.full((h, w), ftzero, np.int32)
    if w > 2:
        s = (a[:, 2:] - a[:, :-2]) * 2 + up[:, 2:] - up[:, :-2] + dn[:, 2:] - dn[:, :-2]
        f[:, 1:-1] = np.clip(s, -ftzero, ftzero) + ftzero
        raw[:, 1:-1] = a[:, 1:-1]
    return f, raw


def _half_pixel_range(p):
    """min / max of (p, (p + left)/2, (p + right)/2) with the row ends replicated."""
    l = np.concatenate([p[:, :1], p[:, :-1]], 1)
    r = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    pl = (p + l) // 2
    pr = (p + r) // 2
    return np.minimum(np.minimum(pl, pr), p), np.maximum(np.maximum(pl, pr), p)


def pixel_cost(L, R, min_disp, num_disp, pre_filter_cap):
    """Stage 2: pix[y, x - maxD', d - minD] (uint8), x in [minX1, maxX1)."""
    h, w = L.shape
    minD, maxD = min_disp, min_disp + num_disp
    minX1, maxX1 = max(maxD, 0), w + min(minD, 0)
    ftzero = max(pre_filter_cap, 15) | 1
    out = np.zeros((h, maxX1 - minX1, num_disp), np.int32)
    xs = np.arange(minX1, maxX1)
    for (pl, pr), scale in zip(zip(prefilter(L, ftzero), prefilter(R, ftzero)), (0, 2)):
        u0, u1 = _half_pixel_range(pl)
        v0, v1 = _half_pixel_range(pr)
        u, U0, U1 = pl[:, xs], u0[:, xs], u1[:, xs]
        for d in range(minD, maxD):
            v, V0, V1 = pr[:, xs - d], v0[:, xs - d], v1[:, xs - d]
            c0 = np.maximum(np.maximum(0, u - V1), V0 - u)
            c1 = np.maximum(np.maximum(0, v - U1), U0 - v)
            out[:, :, d - minD] += np.minimum(c0, c1) >> scale
    return out.astype(np.uint8)


def box_cost(pix, block):
    """Stage 3: C[y, x, d] int16, replicate borders in y (image) and x (computed range)."""
    h, w1, _ = pix.shape
    r = block // 2
    p = pix.astype(np.int32)
    yi = np.clip(np.arange(-r, h + r), 0, h - 1)
    xi = np.clip(np.arange(-r, w1 + r), 0, w1 - 1)
    cs = np.cumsum(np.concatenate([np.zeros((1,) + p.shape[1:], np.int32), p[yi]], 0), 0)
    v = cs[block:] - cs[:-block]
    cs = np.cumsum(np.concatenate([np.zeros((h, 1, p.shape[2]), np.int32), v[:, xi]], 1), 1)
    return (cs[:, block:] - cs[:, :-block]).astype(np.int16)


def _step(C, prevL, prevMin, P1, P2):
    """One step of formula 13 along a path; arrays (..., D) int32."""
    delta = (P2 + prevMin)[..., None]
    pad = np.full(prevL.shape[:-1] + (1,), MAX_COST, np.int32)
    lo = np.concatenate([pad, prevL[..., :-1]], -1) + P1
    hi = np.concatenate([prevL[..., 1:], pad], -1) + P1
    L = C + np.minimum(np.minimum(prevL, lo), np.minimum(hi, delta)) - prevMin[..., None]
    return L, L.min(-1)


def path_costs(C, P1, P2):
    """Stage 4: the five path-cost volumes (int16) of MODE_SGBM."""
    h, w1, D = C.shape
    Ci = C.astype(np.int32)
    out = [np.zeros((h, w1, D), np.int16) for _ in range(5)]
    # horizontal, all rows at once
    for k, xr in ((0, range(w1)), (4, range(w1 - 1, -1, -1))):
        pl = np.zeros((h, D), np.int32)
        pm = np.zeros(h, np.int32)
        for x in xr:
            pl, pm = _step(Ci[:, x], pl, pm, P1, P2)
            out[k][:, x] = pl
    # top-down, all columns at once: predecessor column x-1, x, x+1 of the row above
    for k, sh in ((1, -1), (2, 0), (3, 1)):
        pl = np.zeros((w1, D), np.int32)
        pm = np.zeros(w1, np.int32)
        for y in range(h):
            if sh == -1:
                ql = np.concatenate([np.zeros((1, D), np.int32), pl[:-1]], 0)
                qm = np.concatenate([np.zeros(1, np.int32), pm[:-1]])
            elif sh == 1:
                ql = np.concatenate([pl[1:], np.zeros((1, D), np.int32)], 0)
                qm = np.concatenate([pm[1:], np.zeros(1, np.int32)])
            else:
                ql, qm = pl, pm
            pl, pm = _step(Ci[y], ql, qm, P1, P2)
            out[k][y] = pl
    return out


def aggregate(Ls):
    s4 = np.clip(Ls[0].astype(np.int32) + Ls[1] + Ls[2] + Ls[3], -32768, 32767)
    return np.clip(s4 + Ls[4], -32768, 32767).astype(np.int16)


def _cdiv(a, b):
    """C integer division (truncation towards zero), b > 0."""
    return np.sign(a) * (np.abs(a) // b)


def select_disparity(S, width, min_disp, uniqueness, disp12_max_diff):
    """Stage 5: raw disparity map (int16, 16x) before the median / speckle filters."""
    h, w1, D = S.shape
    minD, maxD = min_disp, min_disp + D
    minX1 = max(maxD, 0)
    inv = (minD - 1) * DISP_SCALE
    d12 = disp12_max_diff if disp12_max_diff > 0 else 1
    Si = S.astype(np.int32)
    best = Si.argmin(-1)                       # first minimum
    minS = Si.min(-1)
    dd = np.arange(D)
    bad = ((Si * (100 - uniqueness) < (minS * 100)[..., None]) & (np.abs(best[..., None] - dd) > 1)).any(-1)
    disp1 = np.full((h, width), inv, np.int32)
    disp2 = np.full((h, width), inv, np.int32)
    cost2 = np.full((h, width), MAX_COST, np.int32)
    rows = np.arange(h)
    for x in range(w1 - 1, -1, -1):
        ok = ~bad[:, x]
        d = best[:, x]
        x2 = x + minX1 - d - minD
        upd = ok & (cost2[rows, x2] > minS[:, x])
        cost2[rows[upd], x2[upd]] = minS[upd, x]
        disp2[rows[upd], x2[upd]] = d[upd] + minD
        inner = (d > 0) & (d < D - 1)
        dm = np.clip(d - 1, 0, D - 1)
        dp = np.clip(d + 1, 0, D - 1)
        sm, s0, sp = Si[rows, x, dm], Si[rows, x, d], Si[rows, x, dp]
        den = np.maximum(sm + sp - 2 * s0, 1)
        sub = d * DISP_SCALE + _cdiv((sm - sp) * DISP_SCALE + den, den * 2)
        val = np.where(inner, sub, d * DISP_SCALE) + minD * DISP_SCALE
        disp1[rows[ok], x + minX1] = val[ok]
    # left-right check
    xs = np.arange(width)[None, :].repeat(h, 0)
    d1 = disp1
    valid = d1 != inv
    lo = d1 >> DISP_SHIFT
    hi = (d1 + DISP_SCALE - 1) >> DISP_SHIFT
    xl, xh = xs - lo, xs - hi
    r2 = rows[:, None].repeat(width, 1)

    def fails(xx, dv):
        inr = (xx >= 0) & (xx < width)
        v = disp2[r2, np.clip(xx, 0, width - 1)]
        return inr & (v >= minD) & (np.abs(v - dv) > d12)
    kill = valid & fails(xl, lo) & fails(xh, hi)
    kill[:, :minX1] = False
    disp1 = np.where(kill, inv, disp1)
    return disp1.astype(np.int16)


def median3(d):
    """medianBlur(disp, disp, 3): replicated border."""
    p = np.pad(d, 1, mode="edge")
    h, w = d.shape
    st = np.stack([p[i:i + h, j:j + w] for i in range(3) for j in range(3)], 0)
    return np.sort(st, 0)[4].astype(d.dtype)


def filter_speckles(d, new_val, max_size, max_diff):
    """cv::filterSpeckles: 4-connected components over edges |a - b| <= max_diff between pixels != new_val;
    components with <= max_size pixels become new_val."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    h, w = d.shape
    v = d.astype(np.int32)
    ok = v != new_val
    idx = np.arange(h * w).reshape(h, w)
    eh = ok[:, :-1] & ok[:, 1:] & (np.abs(v[:, :-1] - v[:, 1:]) <= max_diff)
    ev = ok[:-1] & ok[1:] & (np.abs(v[:-1] - v[1:]) <= max_diff)
    a = np.concatenate([idx[:, :-1][eh], idx[:-1][ev]])
    b = np.concatenate([idx[:, 1:][eh], idx[1:][ev]])
    g = coo_matrix((np.ones(len(a), np.int8), (a, b)), shape=(h * w, h * w))
    _, lab = connected_components(g, directed=False)
    cnt = np.bincount(lab)
    small = (cnt[lab] <= max_size).reshape(h, w) & ok
    out = d.copy()
    out[small] = new_val
    return out


def sgbm_stages(L, R, min_disp=1, num_disp=96, block=7, P1=24, P2=96, disp12_max_diff=0, pre_filter_cap=60,
                uniqueness=0, speckle_window=3000, speckle_range=5):
    """All intermediate stages as a dict (pix, C, S, raw, median, disp)."""
    L = np.asarray(L, np.uint8)
    R = np.asarray(R, np.uint8)
    h, w = L.shape
    block = block if block > 0 else 5
    P1 = P1 if P1 > 0 else 2
    P2 = max(P2 if P2 > 0 else 5, P1 + 1)
    uniq = uniqueness if uniqueness >= 0 else 10
    minD, maxD = min_disp, min_disp + num_disp
    inv = (minD - 1) * DISP_SCALE
    if not (w - maxD > block // 2):
        # cv2 4.13.0 throws here ("input images are too small for your window size and max disparity",
        # stereosgbm.cpp:511); the C ABI returns VO_ERR_BAD_ARG
        raise ValueError("width - (minDisparity + numDisparities) must exceed blockSize / 2")
    if max(maxD, 0) >= w + min(minD, 0):
        z = np.full((h, w), inv, np.int16)
        return dict(raw=z, median=z, disp=z)
    pix = pixel_cost(L, R, min_disp, num_disp, pre_filter_cap)
    C = box_cost(pix, block)
    S = aggregate(path_costs(C, P1, P2))
    raw = select_disparity(S, w, min_disp, uniq, disp12_max_diff)
    med = median3(raw)
    out = med
    if speckle_window > 0:
        out = filter_speckles(med, inv, speckle_window, DISP_SCALE * speckle_range)
    return dict(pix=pix, C=C, S=S, raw=raw, median=med, disp=out)


def sgbm_compute(L, R, **kw):
    return sgbm_stages(L, R, **kw)["disp"]


def rectify_q(fx, fy, cx, cy, baseline, width, height):
    """Q of the reference's stereoRectify call (StereoCV.cpp:224-229: K twice, zero distortion, R = I,
    t = (baseline, 0, 0)) via cv2."""
    import cv2
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    z = np.zeros((4, 1))
    t = np.array([[baseline], [0.0], [0.0]])
    return cv2.stereoRectify(K, z, K, z, (width, height), np.eye(3), t)[4]


def reproject_call_through(disp, Q):
    """reprojectDisparity, StereoCV.cpp:230-247, through cv2: (xyz kept (M,3) float32, flat pixel index (M,))."""
    import cv2
    img3d = cv2.reprojectImageTo3D(np.asarray(disp).astype(np.float32), Q)
    z = img3d[..., 2]
    keep = ~((z > 5) | (z <= 0.01))
    pts = img3d[keep].copy()
    pts[:, 1] *= -1
    return pts, np.flatnonzero(keep.ravel())


def reproject(disp, Q):
    """Restatement of reprojectImageTo3D (CV_32F disparity in, CV_32FC3 out, handleMissingValues = false) + the
    reference's gate: [X Y Z W] = Q * [x y d 1] in double (sum in column order starting from 0), X, Y, Z
    rounded to float, then each divided by the double W and rounded again (cv2 4.13.0 bit-exact; a zero W gives
    inf / nan, which the gate drops)."""
    d = np.asarray(disp).astype(np.float32).astype(np.float64)
    h, w = d.shape
    x = np.arange(w, dtype=np.float64)[None, :].repeat(h, 0)
    y = np.arange(h, dtype=np.float64)[:, None].repeat(w, 1)
    Q = np.asarray(Q, np.float64)
    with np.errstate(all="ignore"):
        hom = [((0.0 + Q[i, 0] * x) + Q[i, 1] * y + Q[i, 2] * d) + Q[i, 3] * 1.0 for i in range(4)]
        xyz = np.stack([(hom[i].astype(np.float32).astype(np.float64) / hom[3]) for i in range(3)],
                       -1).astype(np.float32)
        z = xyz[..., 2]
        keep = ~((z > 5) | (z <= 0.01))
    pts = xyz[keep].copy()
    pts[:, 1] *= -1
    return pts, np.flatnonzero(keep.ravel())
