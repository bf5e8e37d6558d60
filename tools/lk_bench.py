"""Stand-alone LK launch (grid step 5, stereo pair of frame 3) for profiling."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_stereo_slam_b200 import VisualFrontEnd
fe = VisualFrontEnd()
L = fe.synth_render(0, 3, 0); R = fe.synth_render(0, 3, 1)
pts = fe.denseKeypointExtractor(L, 5)
fe.profile_enable(["lk"])
for i in range(5):
    p, st, err = fe.calcOpticalFlowPyrLK(L, R, pts)
w0 = fe.lk_work()
fe.profile_read(reset=True)
for i in range(10):
    p, st, err = fe.calcOpticalFlowPyrLK(L, R, pts)
w1 = fe.lk_work()
l, ms = fe.profile_read()["lk"]
pl, it = (w1[0] - w0[0]) / 10, (w1[1] - w0[1]) / 10
ops = 441 * (30 * pl + 13 * it)
print("points", len(pts), "status", int(st.sum()), "lk ms", ms / l, "point_levels", pl, "iters", it, "Tops/s", ops / (ms / l * 1e-3) / 1e12)
