"""Per-kernel timeline of one frame of the bench workload (both chains)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_stereo_slam_b200 import VisualFrontEnd, _lib
fe = VisualFrontEnd(grid_step=5, pnp_iters=1024, kf_min_inliers=2**31 - 1, ransac_exhaustive=1)
W, H = 1241, 376
d = C.c_void_p()
_lib.check(fe.lib.vo_alloc_dev(fe.h, C.byref(d), C.c_uint64(16 * W * H)))
for i in range(8):
    for eye in (0, 1):
        _lib.check(fe.lib.vo_synth_render_dev(fe.h, 0, i, eye, C.c_void_p(d.value + (2 * i + eye) * W * H)))
Ls = [d.value + (2 * i) * W * H for i in range(8)]
Rs = [d.value + (2 * i + 1) * W * H for i in range(8)]
fe.seq_init(Ls[0], Rs[0], is_device=True)
ANN = os.environ.get("VO_TIMELINE_ANNOUNCE", "1") == "1"
for i in range(1, 6):
    if ANN:
        fe.seq_announce(Ls[i + 1], Rs[i + 1])
    fe.seq_track(Ls[i], Rs[i], is_device=True)
fe.profile_enable("all"); fe.profile_read(reset=True)
import time
if ANN:
    fe.seq_announce(Ls[7], Rs[7])
t0 = time.perf_counter(); fe.seq_track(Ls[6], Rs[6], is_device=True); dt = time.perf_counter() - t0
rows = np.zeros((512, 4), np.float32); n = C.c_int()
_lib.check(fe.lib.vo_debug_timeline(fe.h, rows.ctypes.data_as(C.c_void_p), 512, C.byref(n)))
rows = rows[:n.value]
order = np.argsort(rows[:, 2])
print("wall ms (with event overhead)", dt * 1e3)
for r in rows[order]:
    print("chain %d  %-12s start %7.3f  dur %6.3f  end %7.3f" % (int(r[0]), _lib.KERNELS[int(r[1])], r[2], r[3], r[2] + r[3]))
