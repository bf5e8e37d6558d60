"""LK bit-parity and timing against the live cv2 (harness; run on the GPU box).

    python tools/lk_parity.py [quick]

Every case must come out status-identical and position/err bit-identical to cv2.calcOpticalFlowPyrLK.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np

from oracle import glue
from ros_stereo_slam_b200 import VisualFrontEnd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "vo_golden_v1.npz"))
L0, L1, R0 = g["L0"], g["L1"], g["R0"]


def harsh(img, seed):
    """high-contrast variant: strong gradients push the float window sums past 2^24 (slow path)"""
    rng = np.random.default_rng(seed)
    n = rng.integers(0, 2, img.shape, dtype=np.uint8) * 255
    n = cv2.GaussianBlur(n, (3, 3), 0.7)
    out = np.where(img > 110, n, 255 - n // 3).astype(np.uint8)
    return out


def shift(img, dx, dy):
    M = np.float32([[1, 0, dx], [0, 1, dy]])
    return cv2.warpAffine(img, M, (img.shape[1], img.shape[0]), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)


def run(fe, name, A, B, pts, reps=0):
    p0, st0, e0 = cv2.calcOpticalFlowPyrLK(A, B, pts.reshape(-1, 1, 2), None)
    p0 = p0.reshape(-1, 2); st0 = st0.ravel(); e0 = e0.ravel()
    s0 = fe.lk_slow_paths(); w0 = fe.lk_work()
    p, st, e = fe.calcOpticalFlowPyrLK(A, B, pts)
    s1 = fe.lk_slow_paths(); w1 = fe.lk_work()
    ok = st0 == 1
    d = np.abs(p - p0).max(1)
    ident = bool(np.array_equal(st, st0) and np.array_equal(p[ok], p0[ok]) and np.array_equal(e[ok], e0[ok]))
    msg = "%-28s n %6d  status_eq %s  pos_identical %.6f  max|d| %.3g  err_identical %.6f  slowA %d/%d  slowB %d/%d" % (
        name, len(pts), np.array_equal(st, st0), float(np.mean(d[ok] == 0)), float(d[ok].max()) if ok.any() else 0.0,
        float(np.mean(e[ok] == e0[ok])), s1[0] - s0[0], w1[0] - w0[0], s1[1] - s0[1], w1[1] - w0[1])
    if reps:
        fe.profile_enable(["lk"]); fe.profile_read(reset=True)
        for _ in range(reps):
            fe.calcOpticalFlowPyrLK(A, B, pts)
        l, ms = fe.profile_read(reset=True)["lk"]
        msg += "  lk %.4f ms" % (ms / l)
    print(("OK   " if ident else "FAIL ") + msg, flush=True)
    return ident


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    fe = VisualFrontEnd(max_points=131072)
    allok = True
    extra = np.array([[5, 5], [1236, 371], [0.4, 200.7], [1240.2, 3.3], [620.5, 375.9], [0, 0], [1240, 375], [3.2, 370.1]], np.float32)
    for step in ((9, 5) if quick else (30, 9, 5, 2)):
        pts = np.concatenate([glue.dense_keypoint_extractor(376, 1241, step), extra])
        for nm, B in (("temporal", L1), ("stereo", R0)):
            allok &= run(fe, "step%d %s" % (step, nm), L0, B, pts, reps=5 if step in (5, 2) else 0)
    # fractional positions
    rng = np.random.default_rng(1)
    pts = (rng.random((20000, 2)) * [1241, 376]).astype(np.float32)
    allok &= run(fe, "random20k temporal", L0, L1, pts)
    # high-contrast frames: slow paths
    H0 = harsh(L0, 0)
    for nm, B in (("harsh shift(1.3,0.6)", shift(H0, 1.3, 0.6)), ("harsh shift(6.5,-3.2)", shift(H0, 6.5, -3.2)),
                  ("harsh vs other", harsh(L1, 0))):
        pts = np.concatenate([glue.dense_keypoint_extractor(376, 1241, 9), extra])
        allok &= run(fe, nm, H0, B, pts, reps=3)
    # saturated checkerboard: the largest possible derivatives
    cb = ((np.indices((376, 1241)).sum(0) // 3) % 2 * 255).astype(np.uint8)
    allok &= run(fe, "checker shift(0.8,0.4)", cb, shift(cb, 0.8, 0.4), glue.dense_keypoint_extractor(376, 1241, 9), reps=0)
    allok &= run(fe, "checker vs noise", cb, harsh(L0, 3), glue.dense_keypoint_extractor(376, 1241, 9), reps=0)
    fe.close()
    # 3-channel input (the reference's imread frames): replicated gray, a genuinely coloured pair, a harsh one
    fe3 = VisualFrontEnd(max_points=131072, channels=3)

    def colorize(img):
        f = img.astype(np.float32)
        return np.stack([f, 255.0 - 0.8 * f, 255.0 * (f / 255.0) ** 0.7], -1).round().clip(0, 255).astype(np.uint8)

    for step in ((9,) if quick else (30, 9, 5)):
        pts = np.concatenate([glue.dense_keypoint_extractor(376, 1241, step), extra])
        allok &= run(fe3, "bgr gray3 step%d temporal" % step, cv2.cvtColor(L0, cv2.COLOR_GRAY2BGR), cv2.cvtColor(L1, cv2.COLOR_GRAY2BGR),
                     pts, reps=5 if step == 5 else 0)
        allok &= run(fe3, "bgr color step%d stereo" % step, colorize(L0), colorize(R0), pts, reps=5 if step == 5 else 0)
    Hg, Hgs = cv2.cvtColor(H0, cv2.COLOR_GRAY2BGR), cv2.cvtColor(shift(H0, 1.3, 0.6), cv2.COLOR_GRAY2BGR)
    allok &= run(fe3, "bgr gray3 harsh shift", Hg, Hgs, np.concatenate([glue.dense_keypoint_extractor(376, 1241, 9), extra]), reps=3)
    allok &= run(fe3, "bgr gray3 harsh vs other", Hg, cv2.cvtColor(harsh(L1, 0), cv2.COLOR_GRAY2BGR), glue.dense_keypoint_extractor(376, 1241, 15))
    cbg = cv2.cvtColor(cb, cv2.COLOR_GRAY2BGR)
    allok &= run(fe3, "bgr gray3 checker", cbg, cv2.cvtColor(shift(cb, 0.8, 0.4), cv2.COLOR_GRAY2BGR), glue.dense_keypoint_extractor(376, 1241, 15))
    allok &= run(fe3, "bgr gray3 vs color (mixed)", cv2.cvtColor(L0, cv2.COLOR_GRAY2BGR), colorize(L1), glue.dense_keypoint_extractor(376, 1241, 15))
    Hc = np.stack([H0, harsh(L0, 5), 255 - H0], -1)
    Hs = np.stack([shift(Hc[:, :, c], 1.3, 0.6) for c in range(3)], -1)
    allok &= run(fe3, "bgr harsh shift(1.3,0.6)", Hc, Hs, np.concatenate([glue.dense_keypoint_extractor(376, 1241, 9), extra]), reps=3)
    allok &= run(fe3, "bgr harsh vs other", Hc, np.ascontiguousarray(Hc[:, ::-1]), glue.dense_keypoint_extractor(376, 1241, 15))
    fe3.close()
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
