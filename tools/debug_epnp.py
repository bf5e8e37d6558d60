import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth, glue, cvrng
from ros_stereo_slam_b200 import VisualFrontEnd
hm = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "hostmath", "libhostmath.so"))
fe = VisualFrontEnd()
X, xy, _, _, _ = synth.pnp_stress_case(500, 0.1, 0.3, seed=3)
S = cvrng.sample_list(500, 5, 3)
K4 = np.array([glue.FX, glue.FY, glue.CX, glue.CY])
names = [("us",0,10),("cws",10,22),("alphas",22,42),("mtm",42,186),("ut",186,330),("d",330,342),("l6x10",342,402),("rho",402,408),("rep",408,411),("b1",411,415),("b2",415,419),("R",420,429),("t",429,432)]
for idx in S:
    o = np.ascontiguousarray(X[idx]); i = np.ascontiguousarray(xy[idx])
    a = np.zeros(432); b = np.zeros(432)
    hm.hm_epnp5_dbg(o.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), K4.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p))
    r = fe.lib.vo_debug_epnp(fe.h, o.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    print("sample", idx, "rc", r)
    for nm, s, e in names:
        d = np.abs(a[s:e] - b[s:e])
        print("  %-7s maxabs %.3e  nexact %d/%d" % (nm, d.max(), (d == 0).sum(), e - s))

print("---- production kernel vs debug kernel")
fe2 = VisualFrontEnd(ransac_exhaustive=1)
S = cvrng.sample_list(500, 5, 16)
r = fe2.solvePnPRansac(X, xy, 16, 1.0, 0.99, samples=S)
last = fe2.last_pnp()
import cv2
for h, idx in enumerate(S):
    o = np.ascontiguousarray(X[idx]); i = np.ascontiguousarray(xy[idx])
    b = np.zeros(432)
    fe.lib.vo_debug_epnp(fe.h, o.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    R = b[420:429].reshape(3, 3); t = b[429:432]
    rv, _ = cv2.Rodrigues(R)
    print(h, "prod", last["models"][h], "dbg", rv.ravel(), t)
