"""The C++ host side of the boundary: include/vo_b200.hpp mirrors the reference's member functions
(include/visualSLAM.h:152-169 of the reference tree).  CPU: the header and its test driver compile and
link against libvo_b200.so.  GPU: the driver runs the reference's call sequence through the mirror and
its outputs equal the ctypes path bit for bit and the oracle within the north_star tolerances."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "mirror_main.cpp")
LIBDIR = os.path.join(ROOT, "ros_stereo_slam_b200")


def _cuda_lib_dir():
    for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib"):
        if os.path.exists(os.path.join(d, "libcudart.so")):
            return d
    return None


def _build(tmp):
    exe = os.path.join(tmp, "mirror_main")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lvo_b200", "-Wl,-rpath," + LIBDIR]
    cd = _cuda_lib_dir()
    if cd:
        cmd += ["-L", cd, "-Wl,-rpath," + cd, "-Wl,-rpath-link," + cd]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    from ros_stereo_slam_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (python -m ros_stereo_slam_b200.build)"
    _build(str(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("cn", [1, 3])
def test_cpp_mirror_matches_ctypes_path_and_oracle(tmp_path, cn):
    import cv2
    from gpu_common import golden, make_frontend
    from oracle import glue
    G = golden()
    frames = {k: (G[k] if cn == 1 else cv2.cvtColor(G[k], cv2.COLOR_GRAY2BGR)) for k in ("L0", "R0", "L1", "R1")}
    for k, v in frames.items():
        np.ascontiguousarray(v).tofile(str(tmp_path / (k + ".raw")))
    exe = _build(str(tmp_path))
    r = subprocess.run([exe, str(tmp_path), "1241", "376", str(cn)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr

    def rd(name, dt, cols):
        a = np.fromfile(str(tmp_path / name), dt)
        return a.reshape(-1, cols) if cols > 1 else a

    ref3d, ref2d = rd("ref3d.bin", np.float32, 3), rd("ref2d.bin", np.float32, 2)
    trk2d, trk3d = rd("trk2d.bin", np.float32, 2), rd("trk3d.bin", np.float32, 3)
    inl, pose = rd("inliers.bin", np.int32, 1), rd("pose.bin", np.float64, 1)
    kf2d, kf3d, moved = rd("kf2d.bin", np.float32, 2), rd("kf3d.bin", np.float32, 3), rd("moved.bin", np.float32, 3)
    # (1) the ctypes mirror of the same C ABI gives the same bits
    fe = make_frontend(channels=cn)
    xyz, r2 = fe.stereoTriangulate(frames["L0"], frames["R0"])
    assert np.array_equal(r2, ref2d) and np.array_equal(xyz, ref3d)
    res = fe.PerspectiveNpointEstimation(frames["L0"], frames["L1"], r2, xyz)
    assert np.array_equal(res["trk2d"], trk2d) and np.array_equal(res["trk3d"], trk3d)
    assert np.array_equal(res["inliers"], inl)
    assert np.array_equal(np.r_[res["rvec"], res["tvec"]], pose)
    fe.close()
    # (2) the oracle (reference glue over cv2)
    xyz0, ref0 = glue.stereo_triangulate(frames["L0"], frames["R0"], 30)
    assert np.array_equal(ref2d, ref0)
    assert (np.abs(ref3d - xyz0).max(1) / np.abs(xyz0).max(1)).max() <= 1e-4
    want = glue.perspective_n_point_estimation(frames["L0"], frames["L1"], ref0, xyz0, iters=100)
    assert np.array_equal(inl, want["inliers"])
    assert np.abs(pose[:3] - want["rvec"]).max() <= 1e-4 and np.abs(pose[3:] - want["tvec"]).max() <= 1e-3
    # (3) keyframe insertion: world points = pose * camera points (src/keyFrameManagement.cpp:20-30)
    assert len(kf2d) == len(kf3d) == len(moved) > 100
    assert np.array_equal(kf3d, moved)
    # (5) the loop detector's ORB through vo::ORB (gray frames): keypoints and descriptors == cv2's
    if cn == 1:
        from oracle import orb
        kp = np.fromfile(str(tmp_path / "orb_kps.bin"), np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"),
                                                                   ("response", "f4"), ("octave", "i4")]))
        desc = rd("orb_desc.bin", np.uint8, 32)
        live = orb.detect_and_compute_call_through(frames["L1"], 500)
        assert len(kp) == len(live["xy"]) > 300
        assert np.array_equal(np.c_[kp["x"], kp["y"]], live["xy"]) and np.array_equal(kp["octave"], live["octave"])
        assert np.array_equal(kp["angle"], live["angle"]) and np.array_equal(kp["response"], live["response"])
        assert (desc != live["desc"]).any(1).mean() <= 0.001
    # (4) dense stereo through vo::StereoProcess (BGR frames only): stereoMatch + reprojectDisparity == cv2
    if cn == 3:
        from oracle import sgbm
        disp = rd("disp.bin", np.int16, 1).reshape(376, 1241)
        gl, gr = (cv2.cvtColor(frames[k], cv2.COLOR_BGR2GRAY) for k in ("L0", "R0"))
        assert np.array_equal(disp, sgbm.sgbm_call_through(gl, gr))
        assert "(reference Q: 0)" in r.stdout
        Q = sgbm.rectify_q(718.856, 718.856, 607.1928, 185.2157, -0.5707, 1241, 376)
        pts, idx = sgbm.reproject_call_through(disp, Q)
        assert np.array_equal(rd("cloud.bin", np.float32, 3), pts)
        assert np.array_equal(rd("colors.bin", np.float32, 3), frames["L0"].reshape(-1, 3)[idx].astype(np.float32))
