"""Scalar restatement of OpenCV's pyramidal Lucas-Kanade.  TEST INFRASTRUCTURE ONLY.

Restates what cv::calcOpticalFlowPyrLK does when the reference calls it with
all defaults (src/tracking.cpp:18,52 of the reference tree: winSize 21x21,
maxLevel 3, criteria (COUNT+EPS, 30, 0.01), flags 0, minEigThreshold 1e-4),
for 1-channel images and for the 3-channel BGR images the reference really
passes (imread default, src/keyFrameManagement.cpp:52,64; every window sum then
runs over the channels, minEig is normalised by the window area, err by area*cn):
buildOpticalFlowPyramid (pyrDown 5-tap [1 4 6 4 1], REFLECT_101), the Scharr
derivative of the previous image, and the per-level iteration in 14-bit
fixed-point bilinear arithmetic.

The window sums follow the FLOAT ACCUMULATION ORDER of OpenCV 4.13's SSE-baseline
build of video/lkpyramid.cpp (the build cv2 4.13.0 ships; lkpyramid.cpp is not
a dispatched file, so AVX2/AVX-512 hosts run the same code).  A window row is
W = 21*cn interleaved samples, walked in steps of 8:
  * A sums: sample x < 8*(W//8) goes to float SIMD lane x&3 as
    lane = fl(fl(fx*fy) + lane) (no FMA); the remaining samples go to ONE scalar
    float accumulator as acc = fl(acc + float(int product));
  * b sums: in each step of 8 the int32 pair sums diff[x]*I[x] + diff[x+4]*I[x+4]
    (x&7 < 4) are converted to float and added to lane x&3; the remaining samples
    go through the scalar accumulator as above;
  * at the end  total = fl(scalar + fl(fl(l0 + l2) + fl(l1 + l3))).
Positions, status and err come out BIT-IDENTICAL to cv2 (tests/test_oracle_lk.py
pins that on ~50k tracks, gray and BGR); pyramids and derivatives are bit-exact.
`exact_sums=True` selects the round-1 variant that adds the window terms
exactly in integers (within ~1e-5 px of cv2, not bit-identical).

It also counts the iterations actually executed per (point, level), which is
the figure the LK roofline in DESIGN.md is computed from.
"""
import numpy as np

try:
    import numba
    _njit = numba.njit(cache=True)
except Exception:  # pragma: no cover
    def _njit(f):
        return f

W_BITS = 14
FLT_SCALE = np.float32(1.0 / (1 << 20))
FLT_EPSILON = np.float32(1.1920929e-07)


def reflect101(i, n):
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    i = np.where(i >= n, 2 * (n - 1) - i, i)
    return i


def pyr_down(img):
    """cv::pyrDown for u8, 1 channel: separable [1 4 6 4 1], (sum + 128) >> 8,
    BORDER_REFLECT_101, output ((w+1)/2, (h+1)/2)."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    src = img.astype(np.int32)
    k = (1, 4, 6, 4, 1)
    # horizontal on every source row, at even columns
    cols = [reflect101(2 * np.arange(ow) + d, w) for d in (-2, -1, 0, 1, 2)]
    hor = sum(kk * src[:, c] for kk, c in zip(k, cols))
    rows = [reflect101(2 * np.arange(oh) + d, h) for d in (-2, -1, 0, 1, 2)]
    ver = sum(kk * hor[r, :] for kk, r in zip(k, rows))
    return ((ver + 128) >> 8).astype(np.uint8)


def build_pyramid(img, max_level=3, win=21):
    """Levels 0..L of buildOpticalFlowPyramid (unpadded); stops early when the
    next level would not be larger than the window."""
    levels = [np.ascontiguousarray(img)]
    for _ in range(max_level):
        nxt = pyr_down(levels[-1])
        if nxt.shape[1] <= win or nxt.shape[0] <= win:
            break
        levels.append(nxt)
    return levels


def scharr_deriv(img):
    """calcSharrDeriv: int16 (h, w, 2) = (dx, dy), REFLECT_101 1-px border, gain 32."""
    h, w = img.shape
    s = img.astype(np.int32)
    up = s[reflect101(np.arange(h) - 1, h), :]
    dn = s[reflect101(np.arange(h) + 1, h), :]
    smooth = (up + dn) * 3 + s * 10      # vertical smoothing
    diff = dn - up                        # vertical derivative
    xl = reflect101(np.arange(w) - 1, w)
    xr = reflect101(np.arange(w) + 1, w)
    dx = smooth[:, xr] - smooth[:, xl]
    dy = (diff[:, xl] + diff[:, xr]) * 3 + diff * 10
    return np.stack([dx, dy], -1).astype(np.int16)


def pad_reflect101(img, b):
    return np.pad(img, ((b, b), (b, b)), mode="reflect")


def pad_zero(d, b):
    return np.pad(d, ((b, b), (b, b), (0, 0)), mode="constant")


@_njit
def _cv_round_f32(v):
    # cvRound(float): round half to even
    return int(np.rint(v))


@_njit
def _weights(a, b):
    one = np.float32(1.0)
    s = np.float32(1 << 14)
    iw00 = _cv_round_f32((one - a) * (one - b) * s)
    iw01 = _cv_round_f32(a * (one - b) * s)
    iw10 = _cv_round_f32((one - a) * b * s)
    iw11 = (1 << 14) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


@_njit
def _track_level(Ipad, Dpad, Jpad, rows, cols, level, max_level, win, max_count, eps, min_eig_thr,
                   prev_pts, next_pts, status, err, iters_out):
    """One pyramid level of LKTrackerInvoker for every point, window sums in OpenCV's float order
    (module docstring).  Ipad/Jpad (cn, H, W) are the REFLECT_101-padded (by win) levels, Dpad
    (cn, H, W, 2) the zero-padded derivative; a window row is the cn-interleaved sample sequence."""
    n = prev_pts.shape[0]
    cn = Ipad.shape[0]
    half = np.float32((win - 1) * 0.5)
    scale = np.float32(1.0 / (1 << level))
    W = win * cn
    Iw = np.zeros((win, W), np.int32)
    Ix = np.zeros((win, W), np.int32)
    Iy = np.zeros((win, W), np.int32)
    flt_scale = np.float32(1.0 / (1 << 20))
    nsimd = (W // 8) * 8
    for p in range(n):
        px = np.float32(prev_pts[p, 0] * scale); py = np.float32(prev_pts[p, 1] * scale)
        if level == max_level:
            nx = px; ny = py
        else:
            nx = np.float32(next_pts[p, 0] * np.float32(2.0)); ny = np.float32(next_pts[p, 1] * np.float32(2.0))
        next_pts[p, 0] = nx; next_pts[p, 1] = ny
        px = np.float32(px - half); py = np.float32(py - half)
        ipx = int(np.floor(px)); ipy = int(np.floor(py))
        if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
            if level == 0:
                status[p] = 0; err[p] = 0
            continue
        a = np.float32(px - np.float32(ipx)); b = np.float32(py - np.float32(ipy))
        iw00, iw01, iw10, iw11 = _weights(a, b)
        qA11 = np.zeros(4, np.float32); qA12 = np.zeros(4, np.float32); qA22 = np.zeros(4, np.float32)
        iA11 = np.float32(0); iA12 = np.float32(0); iA22 = np.float32(0)
        for y in range(win):
            yy = y + ipy + win
            for x in range(W):
                c = x % cn; xc = x // cn
                xx = xc + ipx + win
                ival = (int(Ipad[c, yy, xx]) * iw00 + int(Ipad[c, yy, xx + 1]) * iw01
                        + int(Ipad[c, yy + 1, xx]) * iw10 + int(Ipad[c, yy + 1, xx + 1]) * iw11 + (1 << 8)) >> 9
                ixv = (int(Dpad[c, yy, xx, 0]) * iw00 + int(Dpad[c, yy, xx + 1, 0]) * iw01
                       + int(Dpad[c, yy + 1, xx, 0]) * iw10 + int(Dpad[c, yy + 1, xx + 1, 0]) * iw11 + (1 << 13)) >> 14
                iyv = (int(Dpad[c, yy, xx, 1]) * iw00 + int(Dpad[c, yy, xx + 1, 1]) * iw01
                       + int(Dpad[c, yy + 1, xx, 1]) * iw10 + int(Dpad[c, yy + 1, xx + 1, 1]) * iw11 + (1 << 13)) >> 14
                Iw[y, x] = ival; Ix[y, x] = ixv; Iy[y, x] = iyv
            for x in range(nsimd):
                k = x & 3
                fx = np.float32(Ix[y, x]); fy = np.float32(Iy[y, x])
                qA22[k] = np.float32(np.float32(fy * fy) + qA22[k])
                qA12[k] = np.float32(np.float32(fx * fy) + qA12[k])
                qA11[k] = np.float32(np.float32(fx * fx) + qA11[k])
            for x in range(nsimd, W):
                iA11 = np.float32(iA11 + np.float32(Ix[y, x] * Ix[y, x]))
                iA12 = np.float32(iA12 + np.float32(Ix[y, x] * Iy[y, x]))
                iA22 = np.float32(iA22 + np.float32(Iy[y, x] * Iy[y, x]))
        # v_reduce_sum of the SSE build: (l0 + l2) + (l1 + l3)
        iA11 = np.float32(iA11 + np.float32(np.float32(qA11[0] + qA11[2]) + np.float32(qA11[1] + qA11[3])))
        iA12 = np.float32(iA12 + np.float32(np.float32(qA12[0] + qA12[2]) + np.float32(qA12[1] + qA12[3])))
        iA22 = np.float32(iA22 + np.float32(np.float32(qA22[0] + qA22[2]) + np.float32(qA22[1] + qA22[3])))
        A11 = np.float32(iA11 * flt_scale); A12 = np.float32(iA12 * flt_scale); A22 = np.float32(iA22 * flt_scale)
        D = np.float32(np.float32(A11 * A22) - np.float32(A12 * A12))
        dd = np.float32(A11 - A22)
        q = np.float32(np.float32(dd * dd) + np.float32(np.float32(np.float32(4.0) * A12) * A12))
        min_eig = np.float32(np.float32(np.float32(A22 + A11) - np.float32(np.sqrt(q))) / np.float32(2 * win * win))
        if min_eig < min_eig_thr or D < np.float32(1.1920929e-07):
            if level == 0:
                status[p] = 0
            continue
        D = np.float32(np.float32(1.0) / D)
        nx = np.float32(nx - half); ny = np.float32(ny - half)
        pdx = np.float32(0.0); pdy = np.float32(0.0)
        j = 0
        while j < max_count:
            inx = int(np.floor(nx)); iny = int(np.floor(ny))
            if inx < -win or inx >= cols or iny < -win or iny >= rows:
                if level == 0:
                    status[p] = 0
                break
            a = np.float32(nx - np.float32(inx)); b = np.float32(ny - np.float32(iny))
            iw00, iw01, iw10, iw11 = _weights(a, b)
            qb0 = np.zeros(4, np.float32); qb1 = np.zeros(4, np.float32)
            ib1 = np.float32(0); ib2 = np.float32(0)
            for y in range(win):
                yy = y + iny + win
                dif = np.zeros(W, np.int32)
                for x in range(W):
                    c = x % cn; xc = x // cn
                    xx = xc + inx + win
                    jv = (int(Jpad[c, yy, xx]) * iw00 + int(Jpad[c, yy, xx + 1]) * iw01
                          + int(Jpad[c, yy + 1, xx]) * iw10 + int(Jpad[c, yy + 1, xx + 1]) * iw11 + (1 << 8)) >> 9
                    dif[x] = jv - Iw[y, x]
                for x0 in range(0, nsimd, 8):
                    # dotprod pairs (k, k+4)
                    for k in range(4):
                        tx = dif[x0 + k] * Ix[y, x0 + k] + dif[x0 + k + 4] * Ix[y, x0 + k + 4]
                        ty = dif[x0 + k] * Iy[y, x0 + k] + dif[x0 + k + 4] * Iy[y, x0 + k + 4]
                        if k < 2:
                            qb0[2 * k] = np.float32(qb0[2 * k] + np.float32(tx)); qb0[2 * k + 1] = np.float32(qb0[2 * k + 1] + np.float32(ty))
                        else:
                            qb1[2 * (k - 2)] = np.float32(qb1[2 * (k - 2)] + np.float32(tx)); qb1[2 * (k - 2) + 1] = np.float32(qb1[2 * (k - 2) + 1] + np.float32(ty))
                for x in range(nsimd, W):
                    ib1 = np.float32(ib1 + np.float32(dif[x] * Ix[y, x]))
                    ib2 = np.float32(ib2 + np.float32(dif[x] * Iy[y, x]))
            iters_out[p] += 1
            s0 = np.float32(qb0[0] + qb1[0]); s1 = np.float32(qb0[1] + qb1[1]); s2 = np.float32(qb0[2] + qb1[2]); s3 = np.float32(qb0[3] + qb1[3])
            ib1 = np.float32(ib1 + np.float32(s0 + s2)); ib2 = np.float32(ib2 + np.float32(s1 + s3))
            b1 = np.float32(ib1 * flt_scale); b2 = np.float32(ib2 * flt_scale)
            dx = np.float32(np.float32(np.float32(A12 * b2) - np.float32(A22 * b1)) * D)
            dy = np.float32(np.float32(np.float32(A12 * b1) - np.float32(A11 * b2)) * D)
            nx = np.float32(nx + dx); ny = np.float32(ny + dy)
            next_pts[p, 0] = np.float32(nx + half); next_pts[p, 1] = np.float32(ny + half)
            if float(dx) * float(dx) + float(dy) * float(dy) <= eps:
                break
            if j > 0 and abs(float(np.float32(dx + pdx))) < 0.01 and abs(float(np.float32(dy + pdy))) < 0.01:
                next_pts[p, 0] = np.float32(next_pts[p, 0] - np.float32(dx * np.float32(0.5)))
                next_pts[p, 1] = np.float32(next_pts[p, 1] - np.float32(dy * np.float32(0.5)))
                break
            pdx = dx; pdy = dy
            j += 1
        if status[p] != 0 and level == 0:
            fx = np.float32(next_pts[p, 0] - half); fy = np.float32(next_pts[p, 1] - half)
            inx = int(np.floor(fx)); iny = int(np.floor(fy))
            if inx < -win or inx >= cols or iny < -win or iny >= rows:
                status[p] = 0
                continue
            a = np.float32(fx - np.float32(inx)); b = np.float32(fy - np.float32(iny))
            iw00, iw01, iw10, iw11 = _weights(a, b)
            e = 0
            for y in range(win):
                yy = y + iny + win
                for x in range(W):
                    c = x % cn; xc = x // cn
                    xx = xc + inx + win
                    jv = (int(Jpad[c, yy, xx]) * iw00 + int(Jpad[c, yy, xx + 1]) * iw01
                          + int(Jpad[c, yy + 1, xx]) * iw10 + int(Jpad[c, yy + 1, xx + 1]) * iw11 + (1 << 8)) >> 9
                    e += abs(jv - Iw[y, x])
            err[p] = np.float32(np.float32(e) / np.float32(32 * win * cn * win))


@_njit
def _track_level_intsum(Ipad, Dpad, Jpad, rows, cols, level, max_level, win, max_count, eps, min_eig_thr,
                 prev_pts, next_pts, status, err, iters_out):
    """One pyramid level of LKTrackerInvoker for every point.  Ipad/Jpad (cn, H, W) are the
    REFLECT_101-padded (by win) levels, Dpad (cn, H, W, 2) the zero-padded derivative."""
    n = prev_pts.shape[0]
    cn = Ipad.shape[0]
    half = np.float32((win - 1) * 0.5)
    scale = np.float32(1.0 / (1 << level))
    Iw = np.zeros((cn, win, win), np.int32)
    Ix = np.zeros((cn, win, win), np.int32)
    Iy = np.zeros((cn, win, win), np.int32)
    flt_scale = np.float32(1.0 / (1 << 20))
    for p in range(n):
        px = np.float32(prev_pts[p, 0] * scale)
        py = np.float32(prev_pts[p, 1] * scale)
        if level == max_level:
            nx = px
            ny = py
        else:
            nx = np.float32(next_pts[p, 0] * np.float32(2.0))
            ny = np.float32(next_pts[p, 1] * np.float32(2.0))
        next_pts[p, 0] = nx
        next_pts[p, 1] = ny
        px = np.float32(px - half)
        py = np.float32(py - half)
        ipx = int(np.floor(px))
        ipy = int(np.floor(py))
        if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
            if level == 0:
                status[p] = 0
                err[p] = 0
            continue
        a = np.float32(px - np.float32(ipx))
        b = np.float32(py - np.float32(ipy))
        iw00, iw01, iw10, iw11 = _weights(a, b)
        sA11 = 0
        sA12 = 0
        sA22 = 0
        for c in range(cn):
            for y in range(win):
                for x in range(win):
                    yy = y + ipy + win
                    xx = x + ipx + win
                    ival = (int(Ipad[c, yy, xx]) * iw00 + int(Ipad[c, yy, xx + 1]) * iw01
                            + int(Ipad[c, yy + 1, xx]) * iw10 + int(Ipad[c, yy + 1, xx + 1]) * iw11 + (1 << 8)) >> 9
                    ixv = (int(Dpad[c, yy, xx, 0]) * iw00 + int(Dpad[c, yy, xx + 1, 0]) * iw01
                           + int(Dpad[c, yy + 1, xx, 0]) * iw10 + int(Dpad[c, yy + 1, xx + 1, 0]) * iw11 + (1 << 13)) >> 14
                    iyv = (int(Dpad[c, yy, xx, 1]) * iw00 + int(Dpad[c, yy, xx + 1, 1]) * iw01
                           + int(Dpad[c, yy + 1, xx, 1]) * iw10 + int(Dpad[c, yy + 1, xx + 1, 1]) * iw11 + (1 << 13)) >> 14
                    Iw[c, y, x] = ival
                    Ix[c, y, x] = ixv
                    Iy[c, y, x] = iyv
                    sA11 += ixv * ixv
                    sA12 += ixv * iyv
                    sA22 += iyv * iyv
        A11 = np.float32(np.float32(sA11) * flt_scale)
        A12 = np.float32(np.float32(sA12) * flt_scale)
        A22 = np.float32(np.float32(sA22) * flt_scale)
        D = np.float32(np.float32(A11 * A22) - np.float32(A12 * A12))
        dd = np.float32(A11 - A22)
        q = np.float32(np.float32(dd * dd) + np.float32(np.float32(np.float32(4.0) * A12) * A12))
        min_eig = np.float32(np.float32(np.float32(A22 + A11) - np.float32(np.sqrt(q))) / np.float32(2 * win * win))
        if min_eig < min_eig_thr or D < np.float32(1.1920929e-07):
            if level == 0:
                status[p] = 0
            continue
        D = np.float32(np.float32(1.0) / D)
        nx = np.float32(nx - half)
        ny = np.float32(ny - half)
        pdx = np.float32(0.0)
        pdy = np.float32(0.0)
        j = 0
        while j < max_count:
            inx = int(np.floor(nx))
            iny = int(np.floor(ny))
            if inx < -win or inx >= cols or iny < -win or iny >= rows:
                if level == 0:
                    status[p] = 0
                break
            a = np.float32(nx - np.float32(inx))
            b = np.float32(ny - np.float32(iny))
            iw00, iw01, iw10, iw11 = _weights(a, b)
            sb1 = 0
            sb2 = 0
            for c in range(cn):
                for y in range(win):
                    for x in range(win):
                        yy = y + iny + win
                        xx = x + inx + win
                        jv = (int(Jpad[c, yy, xx]) * iw00 + int(Jpad[c, yy, xx + 1]) * iw01
                              + int(Jpad[c, yy + 1, xx]) * iw10 + int(Jpad[c, yy + 1, xx + 1]) * iw11 + (1 << 8)) >> 9
                        diff = jv - Iw[c, y, x]
                        sb1 += diff * Ix[c, y, x]
                        sb2 += diff * Iy[c, y, x]
            iters_out[p] += 1
            b1 = np.float32(np.float32(sb1) * flt_scale)
            b2 = np.float32(np.float32(sb2) * flt_scale)
            dx = np.float32(np.float32(np.float32(A12 * b2) - np.float32(A22 * b1)) * D)
            dy = np.float32(np.float32(np.float32(A12 * b1) - np.float32(A11 * b2)) * D)
            nx = np.float32(nx + dx)
            ny = np.float32(ny + dy)
            next_pts[p, 0] = np.float32(nx + half)
            next_pts[p, 1] = np.float32(ny + half)
            if float(dx) * float(dx) + float(dy) * float(dy) <= eps:
                break
            if j > 0 and abs(float(np.float32(dx + pdx))) < 0.01 and abs(float(np.float32(dy + pdy))) < 0.01:
                next_pts[p, 0] = np.float32(next_pts[p, 0] - np.float32(dx * np.float32(0.5)))
                next_pts[p, 1] = np.float32(next_pts[p, 1] - np.float32(dy * np.float32(0.5)))
                break
            pdx = dx
            pdy = dy
            j += 1
        if status[p] != 0 and level == 0:
            fx = np.float32(next_pts[p, 0] - half)
            fy = np.float32(next_pts[p, 1] - half)
            inx = int(np.floor(fx))
            iny = int(np.floor(fy))
            if inx < -win or inx >= cols or iny < -win or iny >= rows:
                status[p] = 0
                continue
            a = np.float32(fx - np.float32(inx))
            b = np.float32(fy - np.float32(iny))
            iw00, iw01, iw10, iw11 = _weights(a, b)
            e = 0
            for c in range(cn):
                for y in range(win):
                    for x in range(win):
                        yy = y + iny + win
                        xx = x + inx + win
                        jv = (int(Jpad[c, yy, xx]) * iw00 + int(Jpad[c, yy, xx + 1]) * iw01
                              + int(Jpad[c, yy + 1, xx]) * iw10 + int(Jpad[c, yy + 1, xx + 1]) * iw11 + (1 << 8)) >> 9
                        e += abs(jv - Iw[c, y, x])
            err[p] = np.float32(np.float32(e) * np.float32(1.0 / (32 * win * cn * win)))


def calc_optical_flow_pyr_lk(prev_img, next_img, prev_pts, win=21, max_level=3, max_count=30, eps=0.01,
                             min_eig_thr=1e-4, return_iters=False, exact_sums=False):
    """Restated cv2.calcOpticalFlowPyrLK(prev, next, pts, None) for u8 images, H x W or H x W x cn
    (channels are processed as planes: pyrDown, Scharr and the bilinear samples are per channel).
    Returns (next_pts (N,2) f32, status (N,) u8, err (N,) f32[, iters (levels,N)])."""
    prev_img = np.asarray(prev_img)
    next_img = np.asarray(next_img)
    planes_p = [prev_img] if prev_img.ndim == 2 else [np.ascontiguousarray(prev_img[:, :, c]) for c in range(prev_img.shape[2])]
    planes_n = [next_img] if next_img.ndim == 2 else [np.ascontiguousarray(next_img[:, :, c]) for c in range(next_img.shape[2])]
    prev_pts = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(prev_pts)
    max_count = min(max(max_count, 0), 100)
    eps = min(max(eps, 0.0), 10.0)
    eps = eps * eps
    lps = [build_pyramid(pl, max_level, win) for pl in planes_p]
    lns = [build_pyramid(pl, max_level, win) for pl in planes_n]
    L = len(lps[0]) - 1
    next_pts = np.zeros((n, 2), np.float32)
    status = np.ones(n, np.uint8)
    err = np.zeros(n, np.float32)
    iters = np.zeros((L + 1, n), np.int32)
    for level in range(L, -1, -1):
        I = lps[0][level]
        Ipad = np.stack([pad_reflect101(lp[level], win) for lp in lps])
        Dpad = np.stack([pad_zero(scharr_deriv(lp[level]), win) for lp in lps])
        Jpad = np.stack([pad_reflect101(ln[level], win) for ln in lns])
        fn = _track_level_intsum if exact_sums else _track_level
        fn(Ipad, Dpad, Jpad, I.shape[0], I.shape[1], level, L, win, max_count, eps,
                     np.float32(min_eig_thr), prev_pts, next_pts, status, err, iters[level])
    if return_iters:
        return next_pts, status, err, iters
    return next_pts, status, err
