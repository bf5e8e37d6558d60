"""Generates tests/golden/vo_golden_v3.npz: SGBM vectors from live cv2 4.13.0 --
`StereoSGBM::create(1, 96, 7, 24, 96, 0, 60, 0, 3000, 5)->compute` (reference src/StereoCV.cpp:39-53) on the stereo
pair stored in vo_golden_v1.npz, and `reprojectImageTo3D` + the reference's gate (StereoCV.cpp:229-247).  Inputs
come from v1, so only outputs are stored.

    python tests/golden/make_golden_v3.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sgbm, synth  # noqa: E402

CROP = (slice(100, 220), slice(300, 700))      # the small case of the CPU tests


def main():
    g1 = np.load(os.path.join(ROOT, "tests", "golden", "vo_golden_v1.npz"))
    L0, R0 = g1["L0"], g1["R0"]
    out = dict(cv2_version=np.array(cv2.__version__))
    out["sgbm_full"] = sgbm.sgbm_call_through(L0, R0)
    out["sgbm_crop_d32"] = sgbm.sgbm_call_through(L0[CROP].copy(), R0[CROP].copy(), num_disp=32)
    out["sgbm_crop_u10"] = sgbm.sgbm_call_through(L0[CROP].copy(), R0[CROP].copy(), num_disp=48, min_disp=0, block=5,
                                                  uniqueness=10, speckle_window=50, speckle_range=2,
                                                  disp12_max_diff=2)
    # Q as the reference builds it (t = +baseline: every z is negative, nothing passes the gate) and with the
    # translation of a left-to-right rig (t = -baseline), which is what yields points
    for name, bl in (("ref", synth.BASELINE), ("neg", -synth.BASELINE)):
        Q = sgbm.rectify_q(synth.FX, synth.FY, synth.CX, synth.CY, bl, L0.shape[1], L0.shape[0])
        pts, idx = sgbm.reproject_call_through(out["sgbm_full"], Q)
        out[f"Q_{name}"] = Q
        out[f"reproj_{name}_n"] = np.array(len(idx))
        out[f"reproj_{name}_idx_sum"] = np.array(int(idx.astype(np.int64).sum()))
        out[f"reproj_{name}_pts_head"] = pts[:2000].copy()
        out[f"reproj_{name}_pts_sum"] = pts.astype(np.float64).sum(0)
    path = os.path.join(ROOT, "tests", "golden", "vo_golden_v3.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
