"""Generates tests/golden/vo_golden_v4.npz: ORB descriptor-stage vectors from live cv2 4.13.0 --
`cv2.ORB_create().compute` on 1,500 caller-made octave-0 keypoints of the left frame of vo_golden_v1.npz, and the
smoothed image ORB samples from (the float separable filter path, see oracle/orb.py).

    python tests/golden/make_golden_v4.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orb  # noqa: E402


def main():
    g1 = np.load(os.path.join(ROOT, "tests", "golden", "vo_golden_v1.npz"))
    img = g1["L0"]
    rng = np.random.default_rng(11)
    n = 1500
    xy = np.c_[rng.uniform(32, 1208, n), rng.uniform(32, 343, n)].astype(np.float32)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    sm = orb.smooth_call_through(img)
    out = dict(cv2_version=np.array(cv2.__version__), orb_xy=xy, orb_angle=ang,
               orb_desc=orb.describe_call_through(img, xy, ang), orb_smooth_sum=np.array(int(sm.astype(np.int64).sum())),
               orb_smooth_row200=sm[200].copy())
    path = os.path.join(ROOT, "tests", "golden", "vo_golden_v4.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
