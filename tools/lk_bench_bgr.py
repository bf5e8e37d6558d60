"""Stand-alone 3-channel LK launch (grid step 5, temporal pair of replicated-gray frames) for profiling."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_stereo_slam_b200 import VisualFrontEnd
g = VisualFrontEnd()
L = g.synth_render(0, 3, 0); R = g.synth_render(0, 4, 0)
pts = g.denseKeypointExtractor(L, 5)
g.close()
fe = VisualFrontEnd(channels=3)
A = np.ascontiguousarray(np.repeat(L[:, :, None], 3, 2)); B = np.ascontiguousarray(np.repeat(R[:, :, None], 3, 2))
fe.profile_enable(["lk"])
for i in range(3):
    p, st, err = fe.calcOpticalFlowPyrLK(A, B, pts)
w0 = fe.lk_work(); s0 = fe.lk_slow_paths()
fe.profile_read(reset=True)
for i in range(5):
    p, st, err = fe.calcOpticalFlowPyrLK(A, B, pts)
w1 = fe.lk_work(); s1 = fe.lk_slow_paths()
l, ms = fe.profile_read()["lk"]
print("points", len(pts), "status", int(st.sum()), "lk ms", ms / l, "point_levels", (w1[0] - w0[0]) / 5, "iters", (w1[1] - w0[1]) / 5,
      "slowA", (s1[0] - s0[0]) / 5, "slowB", (s1[1] - s0[1]) / 5)
