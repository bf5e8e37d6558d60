"""Timing of BASELINE.json configs 3 and 4 (the stress cases; parity is in test_gpu_stages.py) -- not a test
module (pytest does not collect it); lives under tests/ because it drives the oracle's scene generator.
    python tests/stress_timing.py
  config 3: 1241x376, step-2 grid candidates (115,134) -> ANMS(80,000) -> 4-level 21x21 LK
  config 4: PnP-RANSAC, N = 20,000, 50 % outliers, 4096 hypotheses (early exit and exhaustive), LM refinement
Each next to the same OpenCV call on the host cores of this box."""
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import glue, synth  # noqa: E402
from ros_stereo_slam_b200 import VisualFrontEnd  # noqa: E402


def best(f, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r


def main():
    cores = len(os.sched_getaffinity(0))
    cv2.setNumThreads(cores)
    fe = VisualFrontEnd(max_points=131072)
    sc = synth.Scene(2)
    L0, L1 = sc.render(0, "L"), sc.render(1, "L")
    cand = glue.dense_keypoint_extractor(376, 1241, 2)
    gx = cv2.Sobel(L0, cv2.CV_32F, 1, 0, ksize=3); gy = cv2.Sobel(L0, cv2.CV_32F, 0, 1, ksize=3)
    resp = cv2.boxFilter(gx * gx + gy * gy, -1, (7, 7))[cand[:, 1].astype(int), cand[:, 0].astype(int)].astype(np.float32)
    t_anms, keep = best(lambda: fe.adaptiveNonMaximalSuppresion(cand, resp, 80000), 3)
    pts = cand[keep]
    fe.profile_enable(["lk", "pyramid"])
    fe.calcOpticalFlowPyrLK(L0, L1, pts)
    fe.profile_read(reset=True)
    t_call, _ = best(lambda: fe.calcOpticalFlowPyrLK(L0, L1, pts))
    pr = fe.profile_read(reset=True)
    t_cv, _ = best(lambda: cv2.calcOpticalFlowPyrLK(L0, L1, pts.reshape(-1, 1, 2), None), 3)
    print("config 3: %d candidates -> ANMS keeps %d (%.1f ms incl. transfers); LK kernel %.3f ms (%.1f Mkeypoints/s), "
          "call with host images and points %.2f ms; cv2 on %d cores %.1f ms"
          % (len(cand), len(keep), t_anms, pr["lk"][1] / pr["lk"][0], len(pts) / (pr["lk"][1] / pr["lk"][0]) / 1e3, t_call, cores, t_cv))
    fe.close()
    X, xy, _, _, _ = synth.pnp_stress_case(20000, 0.5, 0.3, seed=3)
    t_cv, r0 = best(lambda: cv2.solvePnPRansac(X.reshape(-1, 1, 3), xy.reshape(-1, 1, 2), glue.K, np.zeros((4, 1)),
                                               None, None, False, 4096, 1.0, 0.99), 3)
    for ex in (0, 1):
        f = VisualFrontEnd(ransac_exhaustive=ex)
        f.solvePnPRansac(X, xy, 4096, 1.0, 0.99)
        f.profile_enable("all"); f.profile_read(reset=True)
        t, r = best(lambda: f.solvePnPRansac(X, xy, 4096, 1.0, 0.99))
        pr = f.profile_read(reset=True)
        n = max(pr["pnp_solve"][0], 1) / 5.0
        print("config 4 (%s): %d hypotheses evaluated, %d inliers, call %.2f ms (solve %.3f, score %.3f, refine %.3f ms "
              "per call); cv2.solvePnPRansac (adaptive stop) %.1f ms"
              % ("exhaustive" if ex else "early exit", len(f.last_pnp()["counts"]), len(r["inliers"]), t,
                 pr["pnp_solve"][1] / 5, pr["pnp_score"][1] / 5, pr["pnp_refine"][1] / 5, t_cv))
        f.close()


if __name__ == "__main__":
    main()
