"""oracle/orb.py (ORB's descriptor stage: the recovered rBRIEF table, the smoothing ORB really applies, the rotated
comparisons) pinned against live cv2 4.13.0 and the golden vectors (tests/golden/vo_golden_v4.npz)."""
import importlib.util
import os

import numpy as np

from oracle import orb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _recover():
    spec = importlib.util.spec_from_file_location("recover_orb_pattern", os.path.join(GOLD, "recover_orb_pattern.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_pattern_table_reproduces_cv2_impulse_maps():
    """every one of the 2 x 37 x 37 single-impulse descriptors cv2 produces is predicted by the embedded table"""
    m = _recover()
    assert orb.PATTERN.shape == (256, 4) and np.abs(orb.PATTERN).max() == 13
    assert np.array_equal(m.predict(orb.PATTERN), m.observe())
    # and a wrong table is noticed
    bad = orb.PATTERN.copy()
    bad[17, 0] += 1
    assert not np.array_equal(m.predict(bad)[:, 17], m.predict(orb.PATTERN)[:, 17])


def test_smoothing_and_descriptors_match_cv2():
    g = np.load(os.path.join(GOLD, "vo_golden_v1.npz"))
    for key in ("L0", "R1"):
        assert np.array_equal(orb.smooth(g[key]), orb.smooth_call_through(g[key]))
    rng = np.random.default_rng(5)
    n = 2000
    xy = np.c_[rng.uniform(32, 1208, n), rng.uniform(32, 343, n)].astype(np.float32)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    ang[:8] = (0, 90, 180, 270, 45, 359.99, 0.01, 135)
    assert np.array_equal(orb.describe(g["L0"], xy, ang), orb.describe_call_through(g["L0"], xy, ang))


def test_golden_v4():
    g = np.load(os.path.join(GOLD, "vo_golden_v1.npz"))
    g4 = np.load(os.path.join(GOLD, "vo_golden_v4.npz"))
    assert np.array_equal(orb.describe(g["L0"], g4["orb_xy"], g4["orb_angle"]), g4["orb_desc"])
    sm = orb.smooth(g["L0"])
    assert int(sm.astype(np.int64).sum()) == int(g4["orb_smooth_sum"]) and np.array_equal(sm[200], g4["orb_smooth_row200"])


def test_orientation_matches_cv2_detect():
    """IC_Angle + fastAtan2: the angle of every octave-0 keypoint cv2.ORB.detect reports, bit for bit"""
    g = np.load(os.path.join(GOLD, "vo_golden_v1.npz"))
    assert orb.umax_table() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    total = 0
    for key in ("L0", "R1"):
        xy, ang = orb.detect_call_through(g[key])
        assert len(xy) > 1500
        assert np.array_equal(orb.ic_angle(g[key], xy), ang)
        total += len(xy)
        # and detectAndCompute's descriptors of those keypoints = describe(ic_angle)
        if key == "L0":
            assert np.array_equal(orb.describe(g[key], xy[:800], orb.ic_angle(g[key], xy[:800])),
                                  orb.describe_call_through(g[key], xy[:800], ang[:800]))
    assert total > 3000


def test_harris_response_matches_cv2_detect():
    g = np.load(os.path.join(GOLD, "vo_golden_v1.npz"))
    for key in ("L0", "R1"):
        xy, _, resp = orb.detect_call_through_full(g[key])
        assert len(xy) > 1500 and np.array_equal(orb.harris_response(g[key], xy), resp)


def test_fast9_matches_cv2():
    g = np.load(os.path.join(GOLD, "vo_golden_v1.npz"))
    for key, thr, nms in (("L0", 20, True), ("R1", 7, True), ("L1", 3, True), ("L0", 5, False)):
        xy, sc, _ = orb.fast9(g[key], thr, nms)
        rxy, rsc = orb.fast9_call_through(g[key], thr, nms)
        assert np.array_equal(xy, rxy)                      # same corners in the same (raster) order
        if nms:                                             # without suppression cv2 does not compute scores
            assert np.array_equal(sc, rsc)
    assert len(orb.fast9(g["L1"], 3)[0]) > 2000


def test_pyramid_and_detect_and_compute_match_cv2():
    """the 8-level INTER_LINEAR_EXACT pyramid, and the whole of ORB::detectAndCompute with the reference's defaults:
    the keypoint SET of every octave (positions, responses, angles) and the descriptors, bit for bit"""
    import cv2
    g = np.load(os.path.join(GOLD, "vo_golden_v1.npz"))
    pyr = orb.build_pyramid(g["L0"])
    assert [p.shape for p in pyr][:3] == [(376, 1241), (313, 1034), (261, 862)]
    prev = g["L0"]
    for lvl in range(1, 8):
        ref = cv2.resize(prev, pyr[lvl].shape[::-1], interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(pyr[lvl], ref)
        prev = ref
    assert orb.features_per_level() == [109, 90, 75, 63, 52, 44, 36, 31]
    for key, nf in (("L0", 500), ("R1", 500), ("L1", 2000)):
        a = orb.detect_and_compute(g[key], nf)
        b = orb.detect_and_compute_call_through(g[key], nf)
        assert len(a["xy"]) == len(b["xy"]) > 300
        for k in ("xy", "octave", "response", "angle", "desc"):
            assert np.array_equal(a[k], b[k]), (key, k)


def test_detect_and_compute_on_small_and_noisy_frames():
    """pure noise (corners almost everywhere, ties in every selection) and frames whose top pyramid levels are smaller
    than the border filter: the levels that still have an interior must be processed, the others contribute nothing"""
    for seed, shape in ((1, (240, 320)), (2, (130, 170))):
        img = np.random.default_rng(seed).integers(0, 256, shape).astype(np.uint8)
        a = orb.detect_and_compute(img, 500)
        b = orb.detect_and_compute_call_through(img, 500)
        assert len(a["xy"]) == len(b["xy"]) > 300
        for k in ("xy", "octave", "response", "angle", "desc"):
            assert np.array_equal(a[k], b[k]), (shape, k)
