"""Single-chain timing of PerspectiveNpointEstimation (vo_pnp_frame) at the bench sizes: fused chain
(default) vs host-driven chain (VO_B200_TEMPORAL_HOST=1)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_stereo_slam_b200 import VisualFrontEnd
fe = VisualFrontEnd(grid_step=5, pnp_iters=1024, ransac_exhaustive=1)
L0 = fe.synth_render(0, 3, 0); R0 = fe.synth_render(0, 3, 1); L1 = fe.synth_render(0, 4, 0)
xyz, ref2d = fe.stereoTriangulate(L0, R0)
for i in range(5):
    r = fe.PerspectiveNpointEstimation(L0, L1, ref2d, xyz)
ts = []
for i in range(30):
    t0 = time.perf_counter(); r = fe.PerspectiveNpointEstimation(L0, L1, ref2d, xyz); ts.append(time.perf_counter() - t0)
print(os.environ.get("VO_B200_TEMPORAL_HOST"), "points", len(ref2d), "tracked", len(r["trk2d"]), "inliers", len(r["inliers"]),
      "median wall ms", np.median(ts) * 1e3, "min", np.min(ts) * 1e3)
