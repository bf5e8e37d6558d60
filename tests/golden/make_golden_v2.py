"""Generates tests/golden/vo_golden_v2.npz: vectors for the rows added after v1 -- 3-channel (BGR) LK and
pyramids (live cv2 4.13.0), SORcloud (oracle/sor.py; PCL is not available, see its header) and BGR2GRAY.
Inputs are derived from the frames stored in vo_golden_v1.npz, so only outputs are stored.

    python tests/golden/make_golden_v2.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import glue, sor  # noqa: E402


def colorize(img):
    """The genuinely coloured test frame used by the BGR tests (same formula in tests/)."""
    f = img.astype(np.float32)
    return np.stack([f, 255.0 - 0.8 * f, 255.0 * (f / 255.0) ** 0.7], -1).round().clip(0, 255).astype(np.uint8)


def main():
    g1 = np.load(os.path.join(ROOT, "tests", "golden", "vo_golden_v1.npz"))
    L0, R0, L1 = g1["L0"], g1["R0"], g1["L1"]
    out = dict(cv2_version=np.array(cv2.__version__))
    pts = glue.dense_keypoint_extractor(376, 1241, 30)
    for kind, (A, B) in (("gray3", (cv2.cvtColor(L0, cv2.COLOR_GRAY2BGR), cv2.cvtColor(L1, cv2.COLOR_GRAY2BGR))),
                         ("color", (colorize(L0), colorize(L1)))):
        p, st, err = cv2.calcOpticalFlowPyrLK(A, B, pts.reshape(-1, 1, 2), None)
        out[f"bgr_{kind}_lk_pts"] = p.reshape(-1, 2)
        out[f"bgr_{kind}_lk_status"] = st.ravel()
        out[f"bgr_{kind}_lk_err"] = err.ravel()
        n, pyr = cv2.buildOpticalFlowPyramid(A, (21, 21), 3, withDerivatives=True)
        out[f"bgr_{kind}_pyr_l3"] = np.ascontiguousarray(pyr[6])
        out[f"bgr_{kind}_pyr_l3_deriv"] = np.ascontiguousarray(pyr[7])
        out[f"bgr_{kind}_pyr_sums"] = np.array([int(pyr[2 * l].astype(np.int64).sum()) for l in range(4)])
        out[f"bgr_{kind}_deriv_abs_sums"] = np.array([int(np.abs(pyr[2 * l + 1].astype(np.int64)).sum()) for l in range(4)])
    out["gray_of_color_sum"] = np.array(int(cv2.cvtColor(colorize(L0), cv2.COLOR_BGR2GRAY).astype(np.int64).sum()))
    out["gray_of_color_row100"] = cv2.cvtColor(colorize(L0), cv2.COLOR_BGR2GRAY)[100].copy()
    # SORcloud on the keyframe cloud of the v1 stereo pair at grid step 9 (cloud = v1's stereo_9_xyz)
    cloud = g1["stereo_9_xyz"]
    keep, dist, thr = sor.sor_cloud(cloud, 200, 0.01, return_all=True)
    out["sor_keep"] = keep
    out["sor_mean_dist"] = dist
    out["sor_threshold"] = np.array(thr)
    path = os.path.join(ROOT, "tests", "golden", "vo_golden_v2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
