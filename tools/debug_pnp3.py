import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import synth, glue, cvrng
from ros_stereo_slam_b200 import VisualFrontEnd
X, xy, _, _, _ = synth.pnp_stress_case(500, 0.1, 0.3, seed=3)
fe = VisualFrontEnd(ransac_exhaustive=1)
Sall = cvrng.sample_list(500, 5, 128)
ref = []
for idx in Sall:
    o = np.ascontiguousarray(X[idx]); i = np.ascontiguousarray(xy[idx]); b = np.zeros(432)
    fe.lib.vo_debug_epnp(fe.h, o.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    ref.append(b[429:432].copy())
ref = np.array(ref)
for H in (16, 17, 24, 32, 33, 64, 100, 128):
    r = fe.solvePnPRansac(X, xy, H, 1.0, 0.99, samples=Sall[:H]); l = fe.last_pnp()
    d = np.abs(l["models"][:, 3:6] - ref[:H]).max(1)
    print("H", H, "n_h", len(l["models"]), "bad", np.nonzero(d > 0)[0])
