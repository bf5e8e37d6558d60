import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros_stereo_slam_b200 import VisualFrontEnd
X, xy, _, _, _ = synth.pnp_stress_case(15000, 0.3, 0.3, seed=3)
fe = VisualFrontEnd(ransac_exhaustive=1, f_exhaustive=1)
fe.profile_enable("all")
for rep in range(3):
    fe.solvePnPRansac(X, xy, 1024, 1.0, 0.99)
fe.profile_read(reset=True)
for rep in range(5):
    r = fe.solvePnPRansac(X, xy, 1024, 1.0, 0.99)
p = fe.profile_read(reset=True)
print("TPB", os.environ.get("VO_SOLVE_TPB"), "pnp_solve ms", p["pnp_solve"][1] / 5, "score", p["pnp_score"][1] / 5, "refine", p["pnp_refine"][1] / 5, "inl", len(r["inliers"]))
