"""oracle/sgbm.py (restatement of cv::StereoSGBM MODE_SGBM + medianBlur + filterSpeckles and of
reprojectImageTo3D, as the reference calls them in src/StereoCV.cpp:39-53,229-247) pinned BIT-IDENTICAL against
live cv2 and against the committed golden vectors (tests/golden/vo_golden_v3.npz)."""
import os

import numpy as np
import pytest

from oracle import sgbm, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CROP = (slice(100, 220), slice(300, 700))


def _pair():
    g = np.load(os.path.join(GOLD, "vo_golden_v1.npz"))
    return g["L0"], g["R0"]


def _noise(kind):
    rng = np.random.default_rng(0)
    if kind == "binary":          # inverted binary noise: costs high enough to saturate S
        a = (rng.integers(0, 2, (90, 200)) * 255).astype(np.uint8)
        return a, 255 - a
    a = rng.integers(0, 256, (90, 200)).astype(np.uint8)
    return a, np.roll(a, -5, 1)


CASES = [
    ("crop", dict(num_disp=32)),
    ("crop", dict(num_disp=32, min_disp=0, uniqueness=10, speckle_window=0)),
    ("crop", dict(num_disp=48, block=5, speckle_window=50, speckle_range=2, disp12_max_diff=2, uniqueness=15)),
    ("crop", dict(num_disp=16, min_disp=-8, uniqueness=5, block=3)),
    ("binary", dict(num_disp=32, block=11, uniqueness=10, speckle_window=0)),
    ("binary", dict(num_disp=32, block=11, speckle_window=0, P1=200, P2=3000)),
    ("shift", dict(num_disp=16, block=9, uniqueness=10, speckle_window=20, speckle_range=1, pre_filter_cap=5)),
    ("shift", dict(num_disp=16, block=1, uniqueness=3, speckle_window=20, speckle_range=1, pre_filter_cap=100,
                   P1=8, P2=32)),
]


@pytest.mark.parametrize("src,kw", CASES)
def test_restatement_matches_cv2(src, kw):
    if src == "crop":
        L, R = _pair()
        L, R = L[CROP].copy(), R[CROP].copy()
    else:
        L, R = _noise(src)
    ref = sgbm.sgbm_call_through(L, R, **kw)
    st = sgbm.sgbm_stages(L, R, **kw)
    assert np.array_equal(ref, st["disp"])
    if src == "binary":
        assert st["S"].max() == 32767          # the saturating case really saturates


def test_golden_crops():
    g = np.load(os.path.join(GOLD, "vo_golden_v3.npz"))
    L, R = _pair()
    L, R = L[CROP].copy(), R[CROP].copy()
    assert np.array_equal(sgbm.sgbm_compute(L, R, num_disp=32), g["sgbm_crop_d32"])
    assert np.array_equal(sgbm.sgbm_compute(L, R, num_disp=48, min_disp=0, block=5, uniqueness=10, speckle_window=50,
                                            speckle_range=2, disp12_max_diff=2), g["sgbm_crop_u10"])


def test_narrow_images():
    import cv2
    L, R = _pair()
    # width - maxD must exceed block / 2: cv2 throws, the restatement raises
    with pytest.raises(cv2.error):
        sgbm.sgbm_call_through(L[:20, :100].copy(), R[:20, :100].copy())
    with pytest.raises(ValueError):
        sgbm.sgbm_compute(L[:20, :100].copy(), R[:20, :100].copy())
    for w in (101, 104, 130):
        a, b = L[:40, :w].copy(), R[:40, :w].copy()
        assert np.array_equal(sgbm.sgbm_compute(a, b), sgbm.sgbm_call_through(a, b))


def test_reproject_matches_cv2_and_golden():
    g = np.load(os.path.join(GOLD, "vo_golden_v3.npz"))
    disp = g["sgbm_full"]
    for name, bl in (("ref", synth.BASELINE), ("neg", -synth.BASELINE)):
        Q = sgbm.rectify_q(synth.FX, synth.FY, synth.CX, synth.CY, bl, 1241, 376)
        assert np.array_equal(Q, g[f"Q_{name}"])
        a, ia = sgbm.reproject_call_through(disp, Q)
        b, ib = sgbm.reproject(disp, Q)
        assert np.array_equal(ia, ib) and np.array_equal(a, b)
        assert len(ib) == int(g[f"reproj_{name}_n"]) and int(ib.astype(np.int64).sum()) == int(g[f"reproj_{name}_idx_sum"])
        assert np.array_equal(b[:2000], g[f"reproj_{name}_pts_head"])
    assert int(g["reproj_ref_n"]) == 0 and int(g["reproj_neg_n"]) > 400000
