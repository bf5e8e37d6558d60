"""Stand-alone launch times of the RANSAC solver kernels on the bench workload's sizes (per-launch CUDA events
of the library's own profiler): pnp_solve_kernel on 1024 EPnP hypotheses, fmat_solve_kernel on 96 seven-point
samples, pnp_refine_kernel, triangulate_kernel on 15k points.
    python tools/solver_bench.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from ros_stereo_slam_b200 import VisualFrontEnd  # noqa: E402


def main():
    X, xy, _, _, _ = synth.pnp_stress_case(10000, 0.2, 0.3, seed=3)
    fe = VisualFrontEnd(ransac_exhaustive=1, pnp_iters=1024)
    for _ in range(3):
        r = fe.solvePnPRansac(X, xy, 1024, 1.0, 0.99)
    fe.profile_enable("all")
    fe.profile_read(reset=True)
    reps = 20
    for _ in range(reps):
        r = fe.solvePnPRansac(X, xy, 1024, 1.0, 0.99)
    pr = fe.profile_read(reset=True)
    print("pnp: %d inliers; per call: solve %.4f ms (%d launches), score %.4f, refine %.4f, select %.4f"
          % (len(r["inliers"]), pr["pnp_solve"][1] / reps, pr["pnp_solve"][0] // reps, pr["pnp_score"][1] / reps,
             pr["pnp_refine"][1] / reps, pr["select"][1] / reps))
    rng = np.random.default_rng(5)
    n = 15000
    p1 = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], 1).astype(np.float32)
    d = rng.uniform(1, 60, n).astype(np.float32)
    p2 = p1.copy()
    p2[:, 0] -= d
    p2[: n // 5] += rng.uniform(-20, 20, (n // 5, 2)).astype(np.float32)
    p2 += rng.normal(0, 0.2, p2.shape).astype(np.float32)
    for _ in range(3):
        fe.findFundamentalMat(p1, p2, 1.0)
    fe.profile_read(reset=True)
    for _ in range(reps):
        F, mask, ni = fe.findFundamentalMat(p1, p2, 1.0)
    pr = fe.profile_read(reset=True)
    print("fmat: %d inliers; per call: solve %.4f ms (%d launches), score %.4f, select %.4f"
          % (ni, pr["fmat_solve"][1] / reps, pr["fmat_solve"][0] // reps, pr["fmat_score"][1] / reps, pr["select"][1] / reps))
    fe.close()


if __name__ == "__main__":
    main()
