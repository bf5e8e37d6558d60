"""Sequence throughput with channels = 3 (the reference's real input: imread's BGR frames, here the synthetic
gray frames replicated to three channels like a KITTI gray PNG read by imread) at the bench sizes."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_stereo_slam_b200 import VisualFrontEnd, _lib
NF = 64
g = VisualFrontEnd()
frames = [(g.synth_render(0, i, 0), g.synth_render(0, i, 1)) for i in range(NF)]
g.close()
for cn in (3, 1):
    fe = VisualFrontEnd(channels=cn, grid_step=5, pnp_iters=1024, kf_min_inliers=2**31 - 1, ransac_exhaustive=1)
    fr = [(np.ascontiguousarray(np.repeat(l[:, :, None], 3, 2)), np.ascontiguousarray(np.repeat(r[:, :, None], 3, 2))) if cn == 3
          else (l, r) for l, r in frames]
    # device-resident copies
    nb = fr[0][0].nbytes
    d = C.c_void_p()
    _lib.check(fe.lib.vo_alloc_dev(fe.h, C.byref(d), C.c_uint64(2 * NF * nb)))
    for i, (l, r) in enumerate(fr):
        _lib.check(fe.lib.vo_memcpy_h2d(fe.h, C.c_void_p(d.value + (2 * i) * nb), l.ctypes.data_as(C.c_void_p), C.c_uint64(nb)))
        _lib.check(fe.lib.vo_memcpy_h2d(fe.h, C.c_void_p(d.value + (2 * i + 1) * nb), r.ctypes.data_as(C.c_void_p), C.c_uint64(nb)))
    stride = 1241 * cn
    fe.seq_init(d.value, d.value + nb, is_device=True, stride=stride)
    for i in range(1, 6):
        fe.seq_track(d.value + 2 * i * nb, d.value + (2 * i + 1) * nb, is_device=True, stride=stride)
    t0 = time.perf_counter()
    for i in range(6, NF):
        res, code = fe.seq_track(d.value + 2 * i * nb, d.value + (2 * i + 1) * nb, is_device=True, stride=stride)
    fe.sync()
    dt = time.perf_counter() - t0
    print("channels", cn, "frames/s", (NF - 6) / dt, "ms/frame", dt / (NF - 6) * 1e3, "last inliers", res.n_inliers, "kf points", res.n_kf_points)
    fe.close()
