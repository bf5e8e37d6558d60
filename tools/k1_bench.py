"""Stand-alone K1 (fused TMA pyramid) + 3-channel LK launches for profiling: a few gray pyramid
builds, then a BGR context: split + per-plane K1 + lk_kernel_c3 on the same stereo pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_stereo_slam_b200 import VisualFrontEnd
fe = VisualFrontEnd()
L = fe.synth_render(0, 3, 0); R = fe.synth_render(0, 3, 1)
pts = fe.denseKeypointExtractor(L, 5)
fe.profile_enable(["pyramid", "lk"])
for i in range(6):
    fe.pyramid_level(L, 0)
fe.profile_read(reset=True)
for i in range(10):
    fe.pyramid_level(L, 0)
l, ms = fe.profile_read(reset=True)["pyramid"]
print("gray  K1 (levels+borders+Scharr) ms per image", ms / l, "launches", l)
fe.close()
fe3 = VisualFrontEnd(channels=3)
L3 = np.stack([L, 255 - L, L // 2 + 60], -1).astype(np.uint8)
R3 = np.stack([R, 255 - R, R // 2 + 60], -1).astype(np.uint8)
fe3.profile_enable(["pyramid", "lk"])
for i in range(3):
    p, st, err = fe3.calcOpticalFlowPyrLK(L3, R3, pts)
fe3.profile_read(reset=True)
w0 = fe3.lk_work()
for i in range(5):
    p, st, err = fe3.calcOpticalFlowPyrLK(L3, R3, pts)
w1 = fe3.lk_work()
pr = fe3.profile_read(reset=True)
pl, it = (w1[0] - w0[0]) / 5, (w1[1] - w0[1]) / 5
ops = 3 * 441 * (30 * pl + 13 * it)
print("bgr   pyramid ms per image (split + K1 x3 planes)", pr["pyramid"][1] / pr["pyramid"][0] * 2, "lk_c3 ms", pr["lk"][1] / pr["lk"][0],
      "points", len(pts), "status", int(st.sum()), "Tops/s", ops / (pr["lk"][1] / pr["lk"][0] * 1e-3) / 1e12)
