import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import synth, glue, cvrng, replay
from ros_stereo_slam_b200 import VisualFrontEnd
X, xy, _, _, _ = synth.pnp_stress_case(500, 0.1, 0.3, seed=3)
r0 = replay.pnp_ransac(X, xy, glue.K, 100, 1.0, 0.99, exhaustive=True)
S = cvrng.sample_list(500, 5, 100)
print("oracle samples == rng list", np.array_equal(S, r0["samples"]))
for ex in (1, 0):
    fe = VisualFrontEnd(ransac_exhaustive=ex)
    for rep in range(3):
        r = fe.solvePnPRansac(X, xy, 100, 1.0, 0.99); l = fe.last_pnp()
        m = len(l["counts"])
        print("ex", ex, "rep", rep, "n_h", m, "counts ok", np.array_equal(l["counts"], r0["counts"][:m]), "best", l["best"], r0["best"], "niters", l["n_iters"], r0["n_iters"], "inl", len(r["inliers"]), len(r0["inliers"]))
    r = fe.solvePnPRansac(X, xy, 100, 1.0, 0.99, samples=S); l = fe.last_pnp(); m = len(l["counts"])
    print("  replay: counts ok", np.array_equal(l["counts"], r0["counts"][:m]), "best", l["best"], "inl", len(r["inliers"]))
    fe.close()
