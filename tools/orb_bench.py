"""ORB descriptor stage timing on the GPU box: vo_orb_describe (host image + keypoints in, descriptors out) against
cv2.ORB.compute with the same caller-made keypoints.   python tools/orb_bench.py [n_keypoints]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cv2
    from oracle import orb
    from ros_stereo_slam_b200 import VisualFrontEnd
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    g = np.load(os.path.join(ROOT, "tests", "golden", "vo_golden_v1.npz"))
    img = g["L0"]
    rng = np.random.default_rng(3)
    xy = np.c_[rng.uniform(32, 1208, n), rng.uniform(32, 343, n)].astype(np.float32)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    fe = VisualFrontEnd()
    d = fe.orbDescribe(img, xy, ang)
    ref = orb.describe_call_through(img, xy, ang)
    print("%d keypoints: %d descriptors differ from cv2 %s" % (n, int((d != ref).any(1).sum()), cv2.__version__))
    ts = []
    for _ in range(200):
        t = time.perf_counter()
        fe.orbDescribe(img, xy, ang)
        ts.append(time.perf_counter() - t)
    print("vo_orb_describe (smoothing + %d descriptors, host in / host out): median %.3f ms" % (n, 1e3 * np.median(ts[50:])))
    cv2.setNumThreads(len(os.sched_getaffinity(0)))
    o = cv2.ORB_create()
    kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, 0, -1) for (x, y), a in zip(xy, ang)]
    tc = []
    for _ in range(10):
        t = time.perf_counter()
        o.compute(img, kps)
        tc.append(time.perf_counter() - t)
    print("cv2.ORB.compute on %d host cores: median %.2f ms" % (len(os.sched_getaffinity(0)), 1e3 * np.median(tc)))
    # the whole of ORB::create()->detectAndCompute, as the loop detector calls it on every frame
    got = fe.orbDetectAndCompute(img, 500)
    want = orb.detect_and_compute_call_through(img, 500)
    same = len(got["xy"]) == len(want["xy"]) and all(np.array_equal(got[k], want[k]) for k in ("xy", "octave", "response", "angle", "desc"))
    print("detectAndCompute(nfeatures 500): %d keypoints, identical to cv2: %s" % (len(got["xy"]), same))
    td = []
    for _ in range(60):
        t = time.perf_counter()
        fe.orbDetectAndCompute(img, 500)
        td.append(time.perf_counter() - t)
    print("vo_orb_detect_and_compute (host image in, keypoints + descriptors out): median %.3f ms" % (1e3 * np.median(td[10:])))
    o5 = cv2.ORB_create(nfeatures=500)
    tc = []
    for _ in range(10):
        t = time.perf_counter()
        o5.detectAndCompute(img, None)
        tc.append(time.perf_counter() - t)
    print("cv2.ORB.detectAndCompute on %d host cores: median %.2f ms" % (len(os.sched_getaffinity(0)), 1e3 * np.median(tc)))
    # a corner-rich frame (about 10,800 FAST corners on level 0): what real road scenes look like to FAST
    tex = cv2.GaussianBlur(np.random.default_rng(7).integers(0, 256, img.shape).astype(np.uint8), (0, 0), 1.2).astype(np.float32)
    rich = np.clip(0.6 * img.astype(np.float32) + 0.9 * (tex - 128) + 50, 0, 255).astype(np.uint8)
    got = fe.orbDetectAndCompute(rich, 500)
    want = orb.detect_and_compute_call_through(rich, 500)
    same = len(got["xy"]) == len(want["xy"]) and all(np.array_equal(got[k], want[k]) for k in ("xy", "octave", "response", "angle", "desc"))
    td = []
    for _ in range(40):
        t = time.perf_counter()
        fe.orbDetectAndCompute(rich, 500)
        td.append(time.perf_counter() - t)
    tc = []
    for _ in range(8):
        t = time.perf_counter()
        o5.detectAndCompute(rich, None)
        tc.append(time.perf_counter() - t)
    print("corner-rich frame: %d keypoints, identical to cv2: %s; vo_orb_detect_and_compute %.3f ms vs cv2 %.2f ms"
          % (len(got["xy"]), same, 1e3 * np.median(td[10:]), 1e3 * np.median(tc)))
    fe.close()


if __name__ == "__main__":
    main()
