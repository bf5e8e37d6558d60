import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth, glue, cvrng
from ros_stereo_slam_b200 import VisualFrontEnd
X, xy, _, _, _ = synth.pnp_stress_case(500, 0.1, 0.3, seed=3)
S = cvrng.sample_list(500, 5, 16)
for ex in (1, 0):
    fe = VisualFrontEnd(ransac_exhaustive=ex)
    r1 = fe.solvePnPRansac(X, xy, 16, 1.0, 0.99, samples=S); l1 = fe.last_pnp()
    r2 = fe.solvePnPRansac(X, xy, 16, 1.0, 0.99); l2 = fe.last_pnp()
    print("exhaustive", ex, "n_h", len(l1["models"]), len(l2["models"]))
    m = min(len(l1["models"]), len(l2["models"]))
    print(" models equal", np.array_equal(l1["models"][:m], l2["models"][:m]), np.abs(l1["models"][:m] - l2["models"][:m]).max(1))
    print(" counts", l1["counts"][:m], l2["counts"][:m])
    fe.close()
