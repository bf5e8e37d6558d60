"""Generates tests/golden/vo_golden_v1.npz from the oracle (cv2 4.13.0 call-through) in this
container.  The reference ships no golden vectors; these pin the
OpenCV behaviour the parity tests compare against, so that the GPU box (which has no
/root/reference) checks against committed numbers as well as against live cv2.

    python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import glue, replay, synth  # noqa: E402


def main():
    sc = synth.Scene(0)
    L0, R0 = sc.stereo_pair(0)
    L1, R1 = sc.stereo_pair(1)
    out = dict(L0=L0, R0=R0, L1=L1, R1=R1, cv2_version=np.array(cv2.__version__))
    # pyramid checksums (levels are reproduced bit-exactly; keep the small levels verbatim)
    n, pyr = cv2.buildOpticalFlowPyramid(L0, (21, 21), 3, withDerivatives=True)
    out["pyr_l3"] = pyr[6][21:-21, 21:-21].copy() if pyr[6].shape[0] > 47 else pyr[6]
    out["pyr_l3_deriv"] = pyr[7]
    out["pyr_sums"] = np.array([int(pyr[2 * l].astype(np.int64).sum()) for l in range(4)])
    out["deriv_abs_sums"] = np.array([int(np.abs(pyr[2 * l + 1].astype(np.int64)).sum()) for l in range(4)])
    for step in (30, 9):
        g = glue.dense_keypoint_extractor(376, 1241, step)
        # raw LK, stereo and temporal
        for name, nxt in (("stereo", R0), ("temporal", L1)):
            p, st, err = cv2.calcOpticalFlowPyrLK(L0, nxt, g.reshape(-1, 1, 2), None)
            out[f"lk_{name}_{step}_pts"] = p.reshape(-1, 2)
            out[f"lk_{name}_{step}_status"] = st.ravel()
            out[f"lk_{name}_{step}_err"] = err.ravel()
        xyz, ref2d = glue.stereo_triangulate(L0, R0, step)
        out[f"stereo_{step}_xyz"] = xyz
        out[f"stereo_{step}_ref2d"] = ref2d
        res = glue.perspective_n_point_estimation(L0, L1, ref2d, xyz, iters=100)
        out[f"pnp_{step}_trk2d"] = res["trk2d"]
        out[f"pnp_{step}_trk3d"] = res["trk3d"]
        out[f"pnp_{step}_ref2d_inl"] = res["ref2d_inl"]
        out[f"pnp_{step}_rvec"] = res["rvec"]
        out[f"pnp_{step}_tvec"] = res["tvec"]
        out[f"pnp_{step}_inliers"] = res["inliers"]
        pose = glue.camera_pose_from_pnp(res["rvec"], res["tvec"])
        out[f"pose_{step}"] = pose
        w3, w2, cam = glue.insert_key_frames(L1, R1, pose, step)
        out[f"kf_{step}_xyz_world"] = w3
        out[f"kf_{step}_ref2d"] = w2
    # PnP stress (reduced N for file size): sample list + per-sample counts
    X, xy, rvec, tvec, _ = synth.pnp_stress_case(4000, 0.5, 0.3, seed=3)
    r = replay.pnp_ransac(X, xy, glue.K, 200, 1.0, 0.99, exhaustive=True)
    out["stress_X"] = X
    out["stress_xy"] = xy
    out["stress_samples"] = r["samples"]
    out["stress_counts"] = r["counts"]
    out["stress_best"] = np.array(r["best"])
    out["stress_niters"] = np.array(r["n_iters"])
    out["stress_inliers"] = r["inliers"]
    out["stress_rvec"] = r["rvec"]
    out["stress_tvec"] = r["tvec"]
    out["stress_hyp"] = np.array([np.concatenate(h) for h in r["hyp"]])
    path = os.path.join(ROOT, "tests", "golden", "vo_golden_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
