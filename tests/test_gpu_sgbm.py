"""GPU parity of the dense-stereo path through the C ABI: vo_sgbm_compute,
vo_stereo_match, vo_reproject_disparity against live cv2 4.13.0 (the reference's StereoSGBM / reprojectImageTo3D
calls, src/StereoCV.cpp:39-53,229-247), the stage-by-stage restatement oracle/sgbm.py and the golden vectors.
Everything here is integer (bit-exact) except the reprojection, which is float and bit-exact as well."""
import os

import numpy as np
import pytest

from gpu_common import golden, make_frontend

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CROP = (slice(100, 220), slice(300, 700))
NAMES = dict(num_disp="num_disparities", min_disp="min_disparity", block="block_size", P1="p1", P2="p2",
             disp12_max_diff="disp12_max_diff", pre_filter_cap="pre_filter_cap", uniqueness="uniqueness_ratio",
             speckle_window="speckle_window_size", speckle_range="speckle_range")


def _abi(kw):
    return {NAMES[k]: v for k, v in kw.items()}


@pytest.fixture(scope="module")
def fe():
    f = make_frontend()
    yield f
    f.close()


def _noise(kind):
    rng = np.random.default_rng(0)
    if kind == "binary":
        a = (rng.integers(0, 2, (90, 200)) * 255).astype(np.uint8)
        return a, 255 - a
    a = rng.integers(0, 256, (90, 200)).astype(np.uint8)
    return a, np.roll(a, -5, 1)


def test_reference_parameters_full_frame(fe):
    from oracle import sgbm
    g = golden()
    g3 = np.load(os.path.join(GOLD, "vo_golden_v3.npz"))
    disp = fe.stereoMatch(g["L0"], g["R0"])
    assert disp.dtype == np.int16 and disp.shape == (376, 1241)
    assert np.array_equal(disp, g3["sgbm_full"])
    assert np.array_equal(disp, sgbm.sgbm_call_through(g["L0"], g["R0"]))
    assert (disp > 0).mean() > 0.85
    t = fe.sgbm_timing()
    assert all(v >= 0 for v in t.values())


def test_stages_match_restatement(fe):
    from oracle import sgbm
    g = golden()
    L, R = g["L0"][CROP].copy(), g["R0"][CROP].copy()
    kw = dict(num_disp=32, uniqueness=10, speckle_window=40, speckle_range=2)
    st = sgbm.sgbm_stages(L, R, **kw)
    out = fe.stereoMatch(L, R, **_abi(kw))
    h, w = L.shape
    f, raw = sgbm.prefilter(L, 61)
    planes = fe.sgbm_stage(3, (4, h, w, 4), np.uint8)
    assert np.array_equal(planes[0, :, :, 0], f) and np.array_equal(planes[1, :, :, 0], raw)
    C = fe.sgbm_stage(0, st["C"].shape, np.int16)
    assert np.array_equal(C, st["C"])
    assert np.array_equal(fe.sgbm_stage(2, (h, w), np.int16), st["raw"])
    assert np.array_equal(out, st["disp"])
    assert np.array_equal(out, sgbm.sgbm_call_through(L, R, **kw))


CASES = [
    ("crop", dict(num_disp=32)),
    ("crop", dict(num_disp=32, min_disp=0, uniqueness=10, speckle_window=0)),
    ("crop", dict(num_disp=48, block=5, speckle_window=50, speckle_range=2, disp12_max_diff=2, uniqueness=15)),
    ("crop", dict(num_disp=16, min_disp=-8, uniqueness=5, block=3)),
    ("crop", dict(num_disp=160, block=3)),                        # 8 disparities per lane
    ("crop", dict(num_disp=256, min_disp=-100, block=9, uniqueness=8)),
    ("binary", dict(num_disp=32, block=11, uniqueness=10, speckle_window=0)),       # S saturates
    ("binary", dict(num_disp=32, block=11, speckle_window=0, P1=200, P2=3000)),
    ("shift", dict(num_disp=16, block=9, uniqueness=10, speckle_window=20, speckle_range=1, pre_filter_cap=5)),
    ("shift", dict(num_disp=16, block=1, uniqueness=3, speckle_window=20, speckle_range=1, pre_filter_cap=100, P1=8,
                   P2=32)),
]


@pytest.mark.parametrize("src,kw", CASES)
def test_parameter_sweep_matches_cv2(fe, src, kw):
    from oracle import sgbm
    if src == "crop":
        g = golden()
        L, R = g["L0"][CROP].copy(), g["R0"][CROP].copy()
    else:
        L, R = _noise(src)
    assert np.array_equal(fe.stereoMatch(L, R, **_abi(kw)), sgbm.sgbm_call_through(L, R, **kw))


def test_other_frames_and_sizes(fe):
    from oracle import sgbm
    g = golden()
    for L, R in ((g["L1"], g["R0"]), (g["L0"][:201, :777].copy(), g["R0"][:201, :777].copy()),
                 (g["L0"][50:90, :101].copy(), g["R0"][50:90, :101].copy())):
        assert np.array_equal(fe.stereoMatch(L, R), sgbm.sgbm_call_through(L, R))
    # back to the big size after a small one (buffers are reused / regrown)
    assert np.array_equal(fe.stereoMatch(g["L0"], g["R0"]), sgbm.sgbm_call_through(g["L0"], g["R0"]))


def test_stereo_match_on_bgr_frames(fe):
    import cv2
    from oracle import sgbm
    g = golden()
    f = g["L0"].astype(np.float32)
    col = lambda a: np.stack([a, 255 - 0.8 * a, 255 * (a / 255) ** 0.7], -1).round().clip(0, 255).astype(np.uint8)
    Lb, Rb = col(f), col(g["R0"].astype(np.float32))
    ref = sgbm.sgbm_call_through(cv2.cvtColor(Lb, cv2.COLOR_BGR2GRAY), cv2.cvtColor(Rb, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(fe.stereoMatch(Lb, Rb), ref)


def test_reproject_disparity(fe):
    from oracle import sgbm, synth
    g = golden()
    g3 = np.load(os.path.join(GOLD, "vo_golden_v3.npz"))
    disp = g3["sgbm_full"]
    for name in ("ref", "neg"):
        Q = g3[f"Q_{name}"]
        pts, idx = fe.reprojectDisparity(disp, Q)
        with np.errstate(all="ignore"):
            a, ia = sgbm.reproject_call_through(disp, Q)
        assert np.array_equal(idx, ia)
        assert np.array_equal(pts, a)                      # bit-exact, float
        assert len(idx) == int(g3[f"reproj_{name}_n"])
        assert np.array_equal(pts[:2000], g3[f"reproj_{name}_pts_head"])
    # device-resident disparity of the last match, and a capacity-limited call
    fe.stereoMatch(g["L0"], g["R0"], download=False)
    pts2, idx2 = fe.reprojectDisparity(None, g3["Q_neg"], shape=(376, 1241))
    assert np.array_equal(idx2, idx) and np.array_equal(pts2, pts)
    pts3, idx3 = fe.reprojectDisparity(disp, g3["Q_neg"], cap=1000)
    assert len(idx3) == 1000 and np.array_equal(idx3, idx[:1000]) and np.array_equal(pts3, pts[:1000])


def test_rejected_arguments(fe):
    from ros_stereo_slam_b200 import VoError
    g = golden()
    L, R = g["L0"], g["R0"]
    with pytest.raises(VoError):          # cv2 throws for this size (width - maxD <= block / 2)
        fe.stereoMatch(L[:20, :100].copy(), R[:20, :100].copy())
    for kw in (dict(num_disparities=40), dict(num_disparities=272), dict(block_size=13), dict(block_size=4),
               dict(block_size=11, p2=30000)):
        with pytest.raises(VoError):
            fe.stereoMatch(L, R, **kw)


def test_random_parameter_fuzz_matches_cv2(fe):
    """40 random (size, parameter) draws inside the supported range, random texture with a random true disparity
    field: the disparity must equal cv2's bit for bit every time."""
    import cv2
    from oracle import sgbm
    rng = np.random.default_rng(2024)
    done = 0
    while done < 40:
        D = int(rng.choice([16, 32, 48, 64, 96, 128, 160, 256]))
        min_d = int(rng.integers(-24, 24))
        block = int(rng.choice([1, 3, 5, 7, 9, 11]))
        cap = int(rng.integers(0, 127))
        P1 = int(rng.integers(1, 200))
        P2 = int(rng.integers(P1 + 1, 2000))
        ftzero = max(cap, 15) | 1
        if block * block * (2 * ftzero + 63) + P2 > 32767:
            continue
        w = int(rng.integers(max(min_d + D, 0) + block // 2 + 2, max(min_d + D, 0) + 260))
        h = int(rng.integers(8, 70))
        kw = dict(num_disp=D, min_disp=min_d, block=block, P1=P1, P2=P2, pre_filter_cap=cap,
                  disp12_max_diff=int(rng.integers(-1, 4)), uniqueness=int(rng.integers(0, 25)),
                  speckle_window=int(rng.choice([0, 10, 60, 400])), speckle_range=int(rng.integers(1, 5)))
        tex = cv2.GaussianBlur(rng.integers(0, 256, (h, w + 300)).astype(np.uint8), (0, 0), float(rng.uniform(0.6, 2.5)))
        shift = int(rng.integers(max(min_d, -20), max(min_d, -20) + min(D, 40)))
        L = np.ascontiguousarray(tex[:, 150:150 + w])
        R = np.ascontiguousarray(tex[:, 150 + shift:150 + shift + w])
        R[h // 2:] = np.roll(R[h // 2:], 3, 1)                     # a second disparity layer
        try:
            want = sgbm.sgbm_call_through(L, R, **kw)
        except cv2.error:
            continue
        got = fe.stereoMatch(L, R, **_abi(kw))
        assert np.array_equal(got, want), (kw, w, h)
        done += 1


def test_full_hd_frame(fe):
    """A 1920x1080 pair (the golden pair upscaled): grids, strides and the 2.3 GB of volumes at a size the reference's
    KITTI frames never reach."""
    import cv2
    from oracle import sgbm
    g = golden()
    L = cv2.resize(g["L0"], (1920, 1080), interpolation=cv2.INTER_LINEAR)
    R = cv2.resize(g["R0"], (1920, 1080), interpolation=cv2.INTER_LINEAR)
    want = sgbm.sgbm_call_through(L, R)
    got = fe.stereoMatch(L, R)
    assert np.array_equal(got, want)
    assert (want > 0).mean() > 0.3
    # and back to the KITTI size with the grown buffers
    assert np.array_equal(fe.stereoMatch(g["L0"], g["R0"]), sgbm.sgbm_call_through(g["L0"], g["R0"]))
