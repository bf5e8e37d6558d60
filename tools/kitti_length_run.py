"""KITTI-00-length run (4,540 tracked frames) with the REFERENCE'S OWN settings: grid step 30 (440 keypoints,
src/triangulation.cpp:89), PnP 100 iterations, keyframe when inliers < 200 (src/VisualSLAM.cpp:120), OpenCV's adaptive
RANSAC stop.  Frames are rendered on the fly by the harness kernel into a small ring of device buffers.  Reports
frames/s, keyframes, the smallest point counts that reached F-RANSAC / PnP (the small-N estimators of OpenCV start below
15 / 6 points) and whether the sequence ever hit the reference's SHUTDOWN path.

    python tools/kitti_length_run.py [frames] [grid_step]
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ros_stereo_slam_b200 import VisualFrontEnd, _lib

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4540
step = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fe = VisualFrontEnd(grid_step=step, pnp_iters=100, kf_min_inliers=200, ransac_exhaustive=0)
W, H, RING = 1241, 376, 8
d = C.c_void_p()
_lib.check(fe.lib.vo_alloc_dev(fe.h, C.byref(d), C.c_uint64(2 * RING * W * H)))
ptr = lambda i, e: d.value + (2 * (i % RING) + e) * W * H


def render(i):
    for e in (0, 1):
        _lib.check(fe.lib.vo_synth_render_dev(fe.h, 0, i, e, C.c_void_p(ptr(i, e))))


render(0)
fe.seq_init(ptr(0, 0), ptr(0, 1), is_device=True)
min_in, min_trk, min_inl, kfs, shutdown = 10 ** 9, 10 ** 9, 10 ** 9, 0, None
t0 = time.perf_counter()
for i in range(1, n_frames + 1):
    render(i)
    res, code = fe.seq_track(ptr(i, 0), ptr(i, 1), is_device=True)
    min_in, min_trk, min_inl = min(min_in, res.n_lk_in), min(min_trk, res.n_tracked), min(min_inl, res.n_inliers)
    kfs += res.keyframe
    if code != 0:
        shutdown = i
        break
fe.sync()
dt = time.perf_counter() - t0
print("frames %d  grid step %d  frames/s %.1f (incl. rendering)  keyframes %d  min points into LK %d  min tracked (into PnP) %d  "
      "min PnP inliers %d  shutdown at frame %s" % (i, step, i / dt, kfs, min_in, min_trk, min_inl, shutdown))
fe.close()
