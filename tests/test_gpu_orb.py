"""GPU parity of the ORB descriptor stage (first step) through the C ABI: vo_orb_smooth and
vo_orb_describe against the golden vectors (cv2 4.13.0 in the build container), oracle/orb.py and live cv2."""
import os

import numpy as np
import pytest

from gpu_common import golden, make_frontend

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fe():
    f = make_frontend()
    yield f
    f.close()


def test_smoothing_is_the_one_orb_applies(fe):
    from oracle import orb
    g = golden()
    g4 = np.load(os.path.join(GOLD, "vo_golden_v4.npz"))
    for key in ("L0", "R1"):
        sm = fe.orbSmooth(g[key])
        assert np.array_equal(sm, orb.smooth(g[key]))                  # the restatement, bit for bit
        # live cv2: identical here; a cv2 dispatched to another instruction set may differ in isolated pixels
        assert (sm != orb.smooth_call_through(g[key])).sum() <= 4
    sm = fe.orbSmooth(g["L0"])
    assert int(sm.astype(np.int64).sum()) == int(g4["orb_smooth_sum"]) and np.array_equal(sm[200], g4["orb_smooth_row200"])
    odd = g["L0"][:77, :131].copy()                                     # odd size, strided input
    assert np.array_equal(fe.orbSmooth(odd), orb.smooth(odd))


def test_descriptors_match_golden_restatement_and_cv2(fe):
    from oracle import orb
    g = golden()
    g4 = np.load(os.path.join(GOLD, "vo_golden_v4.npz"))
    d = fe.orbDescribe(g["L0"], g4["orb_xy"], g4["orb_angle"])
    assert d.shape == (1500, 32)
    assert np.array_equal(d, g4["orb_desc"])                           # cv2.ORB.compute, generated in the build container
    rng = np.random.default_rng(9)
    n = 5000
    xy = np.c_[rng.uniform(32, 1208, n), rng.uniform(32, 343, n)].astype(np.float32)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    ang[:8] = (0, 90, 180, 270, 45, 359.99, 0.01, 135)
    d = fe.orbDescribe(g["L1"], xy, ang)
    assert np.array_equal(d, orb.describe(g["L1"], xy, ang))
    live = orb.describe_call_through(g["L1"], xy, ang)
    assert (d != live).any(1).mean() <= 0.001                          # see test_smoothing: identical unless cv2's blur differs


def test_rejected_keypoints(fe):
    from ros_stereo_slam_b200 import VoError
    g = golden()
    with pytest.raises(VoError):
        fe.orbDescribe(g["L0"], np.array([[10.0, 100.0]], np.float32), np.zeros(1, np.float32))
    assert fe.orbDescribe(g["L0"], np.zeros((0, 2), np.float32), np.zeros(0, np.float32)).shape == (0, 32)


def test_orientation_and_detect_and_compute_descriptors(fe):
    """IC_Angle on the GPU = the angles cv2.ORB.detect reports; describe(angle_deg=None) = detectAndCompute's descriptors"""
    from oracle import orb
    g = golden()
    for key in ("L0", "R1"):
        xy, ang = orb.detect_call_through(g[key])
        got = fe.orbAngles(g[key], xy)
        assert np.array_equal(got, orb.ic_angle(g[key], xy))
        assert np.array_equal(got, ang)                                # live cv2, bit for bit (integer moments + fastAtan2)
    xy, ang = orb.detect_call_through(g["L0"])
    d = fe.orbDescribe(g["L0"], xy)                                    # angles computed on the device
    assert np.array_equal(d, orb.describe(g["L0"], xy, ang))
    assert (d != orb.describe_call_through(g["L0"], xy, ang)).any(1).mean() <= 0.001


def test_harris_response(fe):
    from oracle import orb
    g = golden()
    for key in ("L0", "R1"):
        xy, _, resp = orb.detect_call_through_full(g[key])
        got = fe.orbHarris(g[key], xy)
        assert np.array_equal(got, orb.harris_response(g[key], xy)) and np.array_equal(got, resp)


def test_fast9_matches_cv2(fe):
    from oracle import orb
    g = golden()
    for key, thr, nms in (("L0", 20, True), ("R1", 7, True), ("L1", 3, True), ("L0", 5, False), ("L0", 1, True)):
        xy, sc = fe.fast9(g[key], thr, nms)
        rxy, rsc = orb.fast9_call_through(g[key], thr, nms)
        assert np.array_equal(xy, rxy), (key, thr, nms, len(xy), len(rxy))
        oxy, osc, _ = orb.fast9(g[key], thr, nms)
        assert np.array_equal(xy, oxy) and np.array_equal(sc, osc)
        if nms:
            assert np.array_equal(sc, rsc)
    odd = g["L1"][:99, :203].copy()
    xy, sc = fe.fast9(odd, 4)
    rxy, rsc = orb.fast9_call_through(odd, 4)
    assert np.array_equal(xy, rxy) and np.array_equal(sc, rsc)
    # the level-0 front of ORB's detector: FAST -> Harris ranking -> orientation -> descriptors, all on the device
    xy, _ = fe.fast9(g["L1"], 20)
    inside = (xy[:, 0] >= 31) & (xy[:, 0] < 1241 - 31) & (xy[:, 1] >= 31) & (xy[:, 1] < 376 - 31)   # edgeThreshold
    xy = xy[inside]
    assert np.array_equal(fe.orbHarris(g["L1"], xy), orb.harris_response(g["L1"], xy))
    assert np.array_equal(fe.orbDescribe(g["L1"], xy), orb.describe(g["L1"], xy, orb.ic_angle(g["L1"], xy)))


def test_detect_and_compute_matches_cv2(fe):
    """the whole of ORB::detectAndCompute on the device: keypoint sets per octave (position, response, angle) and
    descriptors equal cv2's and the restatement's"""
    from oracle import orb
    g = golden()
    for key, nf in (("L0", 500), ("R1", 500), ("L1", 2000)):
        got = fe.orbDetectAndCompute(g[key], nf)
        want = orb.detect_and_compute(g[key], nf)
        live = orb.detect_and_compute_call_through(g[key], nf)
        assert len(got["xy"]) == len(want["xy"]) == len(live["xy"]) > 300
        for k in ("xy", "octave", "response", "angle", "desc"):
            assert np.array_equal(got[k], want[k]), (key, k)
        for k in ("xy", "octave", "response", "angle"):
            assert np.array_equal(got[k], live[k]), (key, k)
        assert (got["desc"] != live["desc"]).any(1).mean() <= 0.001     # exact unless cv2's blur dispatch differs


def _textured(img):
    """a corner-rich variant of the golden frame (about 10,800 FAST corners on level 0 instead of 35)"""
    import cv2
    rng = np.random.default_rng(7)
    tex = cv2.GaussianBlur(rng.integers(0, 256, img.shape).astype(np.uint8), (0, 0), 1.2).astype(np.float32)
    return np.clip(0.6 * img.astype(np.float32) + 0.9 * (tex - 128) + 50, 0, 255).astype(np.uint8)


def test_detect_and_compute_on_a_corner_rich_frame(fe):
    from oracle import orb
    img = _textured(golden()["L0"])
    assert len(orb.fast9_call_through(img, 20)[0]) > 8000
    for nf in (500, 1500):
        got = fe.orbDetectAndCompute(img, nf)
        live = orb.detect_and_compute_call_through(img, nf)
        assert len(got["xy"]) == len(live["xy"]) >= nf
        for k in ("xy", "octave", "response", "angle"):
            assert np.array_equal(got[k], live[k]), (nf, k)
        assert (got["desc"] != live["desc"]).any(1).mean() <= 0.001


def test_detect_and_compute_other_sizes(fe):
    """640 x 480 and an odd 517 x 389 frame: other level sizes, other resize tables, buffers regrown and reused"""
    import cv2
    from oracle import orb
    base = _textured(golden()["L0"])
    noise = np.random.default_rng(1).integers(0, 256, (240, 320)).astype(np.uint8)     # corners almost everywhere
    small = np.random.default_rng(2).integers(0, 256, (130, 170)).astype(np.uint8)     # top levels smaller than the border
    for size in ((640, 480), (517, 389), (1241, 376), noise, small):
        img = size if isinstance(size, np.ndarray) else cv2.resize(base, size, interpolation=cv2.INTER_AREA)
        got = fe.orbDetectAndCompute(img, 500)
        live = orb.detect_and_compute_call_through(img, 500)
        assert len(got["xy"]) == len(live["xy"]) > 200, img.shape
        for k in ("xy", "octave", "response", "angle"):
            assert np.array_equal(got[k], live[k]), (img.shape, k)
        assert (got["desc"] != live["desc"]).any(1).mean() <= 0.001


def test_detect_and_compute_keeps_ties_like_cv2(fe):
    """a lattice of single bright pixels: thousands of corners share one FAST score and one Harris response, so both
    retainBest selections keep far more than their quota (ties stay) and the packed result exceeds the first transfer"""
    from oracle import orb
    yy, xx = np.mgrid[0:480, 0:640]
    img = (((xx % 10) == 3) & ((yy % 10) == 3)).astype(np.uint8) * 200 + 20
    got = fe.orbDetectAndCompute(img, 100)
    live = orb.detect_and_compute_call_through(img, 100)
    assert len(got["xy"]) == len(live["xy"]) > 2000, len(live["xy"])
    for k in ("xy", "octave", "response", "angle"):
        assert np.array_equal(got[k], live[k]), k
    assert (got["desc"] != live["desc"]).any(1).mean() <= 0.001
