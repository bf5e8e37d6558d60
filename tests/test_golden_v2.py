"""Committed vectors for the rows added after v1 (tests/golden/make_golden_v2.py): 3-channel LK / pyramids
(cv2 4.13.0), SORcloud (oracle/sor.py, parity unpinned) and BGR2GRAY.  CPU: the oracle reproduces them.
GPU: the CUDA path reproduces them through the C ABI."""
import os

import cv2
import numpy as np
import pytest

from oracle import glue, lk as olk, sor as osor

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    return (np.load(os.path.join(HERE, "golden", "vo_golden_v1.npz")),
            np.load(os.path.join(HERE, "golden", "vo_golden_v2.npz")))


def _colorize(img):
    f = img.astype(np.float32)
    return np.stack([f, 255.0 - 0.8 * f, 255.0 * (f / 255.0) ** 0.7], -1).round().clip(0, 255).astype(np.uint8)


def _frames(g1, kind):
    if kind == "gray3":
        return cv2.cvtColor(g1["L0"], cv2.COLOR_GRAY2BGR), cv2.cvtColor(g1["L1"], cv2.COLOR_GRAY2BGR)
    return _colorize(g1["L0"]), _colorize(g1["L1"])


@pytest.mark.parametrize("kind", ["gray3", "color"])
def test_oracle_reproduces_bgr_vectors(kind):
    g1, g2 = _load()
    A, B = _frames(g1, kind)
    pts = glue.dense_keypoint_extractor(376, 1241, 30)
    p, st, err = olk.calc_optical_flow_pyr_lk(A, B, pts)
    assert np.array_equal(st, g2[f"bgr_{kind}_lk_status"])
    ok = st == 1
    assert np.array_equal(p[ok], g2[f"bgr_{kind}_lk_pts"][ok])        # cv2's bits
    assert np.array_equal(err[ok], g2[f"bgr_{kind}_lk_err"][ok])
    for c in range(3):
        lv = olk.build_pyramid(np.ascontiguousarray(A[:, :, c]), 3, 21)
        assert np.array_equal(lv[3], g2[f"bgr_{kind}_pyr_l3"][:, :, c])
        assert np.array_equal(olk.scharr_deriv(lv[3]), g2[f"bgr_{kind}_pyr_l3_deriv"][:, :, 2 * c:2 * c + 2])
    assert int(glue.bgr_to_gray_restated(_colorize(g1["L0"])).astype(np.int64).sum()) == int(g2["gray_of_color_sum"])


def test_oracle_reproduces_sor_vectors():
    g1, g2 = _load()
    keep, dist, thr = osor.sor_cloud(g1["stereo_9_xyz"], 200, 0.01, return_all=True)
    assert np.array_equal(keep, g2["sor_keep"]) and np.array_equal(dist, g2["sor_mean_dist"])
    assert thr == float(g2["sor_threshold"])


@pytest.mark.gpu
def test_cuda_reproduces_v2_vectors():
    from gpu_common import make_frontend
    g1, g2 = _load()
    fe3 = make_frontend(channels=3)
    pts = glue.dense_keypoint_extractor(376, 1241, 30)
    for kind in ("gray3", "color"):
        A, B = _frames(g1, kind)
        p, st, err = fe3.calcOpticalFlowPyrLK(A, B, pts)
        assert np.array_equal(st, g2[f"bgr_{kind}_lk_status"])
        ok = st == 1
        assert np.array_equal(p[ok], g2[f"bgr_{kind}_lk_pts"][ok])        # cv2's bits
        assert np.array_equal(err[ok], g2[f"bgr_{kind}_lk_err"][ok])
        lv, dv = fe3.pyramid_level(A, 3)
        assert np.array_equal(lv, g2[f"bgr_{kind}_pyr_l3"]) and np.array_equal(dv, g2[f"bgr_{kind}_pyr_l3_deriv"])
        for l in range(4):
            lv, dv = fe3.pyramid_level(A, l)
            assert int(lv.astype(np.int64).sum()) == int(g2[f"bgr_{kind}_pyr_sums"][l])
            assert int(np.abs(dv.astype(np.int64)).sum()) == int(g2[f"bgr_{kind}_deriv_abs_sums"][l])
    fe3.close()
    fe = make_frontend()
    gray = fe.cvtColorBGR2GRAY(_colorize(g1["L0"]))
    assert int(gray.astype(np.int64).sum()) == int(g2["gray_of_color_sum"])
    assert np.array_equal(gray[100], g2["gray_of_color_row100"])
    _pts, _c, idx, md = fe.SORcloud(g1["stereo_9_xyz"], None, 200, 0.01, return_distances=True)
    assert np.array_equal(idx, g2["sor_keep"]) and np.array_equal(md, g2["sor_mean_dist"])
    fe.close()
