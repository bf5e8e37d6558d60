"""SORcloud timing: keyframe cloud of the bench workload (grid step 5) through vo_sor_cloud
vs scipy's cKDTree kNN on the host cores (PCL itself is not available in this image)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_stereo_slam_b200 import VisualFrontEnd
fe = VisualFrontEnd(grid_step=5)
L = fe.synth_render(0, 3, 0); R = fe.synth_render(0, 3, 1)
xyz, _ = fe.stereoTriangulate(L, R)
fe.profile_enable(["misc"])
for i in range(3):
    pts, _c = fe.SORcloud(xyz)
fe.profile_read(reset=True)
t0 = time.perf_counter()
for i in range(10):
    pts, _c = fe.SORcloud(xyz)
wall = (time.perf_counter() - t0) / 10
l, ms = fe.profile_read()["misc"]
print("cloud", len(xyz), "kept", len(pts), "kernels per call", l / 10, "kernel ms per call (without the CUB sort)", ms / 10,
      "call wall ms", wall * 1e3)
from scipy.spatial import cKDTree
p64 = xyz.astype(np.float64)
t0 = time.perf_counter()
tree = cKDTree(p64)
d, _ = tree.query(p64, 201, workers=-1)
print("scipy cKDTree build + kNN(201), all cores: ms", (time.perf_counter() - t0) * 1e3)
