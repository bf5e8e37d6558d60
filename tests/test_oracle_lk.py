"""Pins oracle.lk (scalar restatement, window sums in OpenCV's float order) against cv2: pyramids and
Scharr derivatives bit-exact; tracked positions, status and err BIT-IDENTICAL, gray and BGR."""
import numpy as np
import cv2
import pytest

from oracle import lk, synth, glue


@pytest.fixture(scope="module")
def frames():
    sc = synth.Scene(0)
    return sc.render(0, "L"), sc.render(1, "L"), sc.render(0, "R")


def test_pyramid_and_scharr_bit_exact(frames):
    L0, _, _ = frames
    n, pyr = cv2.buildOpticalFlowPyramid(L0, (21, 21), 3, withDerivatives=True)
    assert n == 3
    mine = lk.build_pyramid(L0, 3, 21)
    assert [m.shape for m in mine] == [(376, 1241), (188, 621), (94, 311), (47, 156)]
    for l in range(4):
        assert np.array_equal(pyr[2 * l], mine[l])
        assert np.array_equal(pyr[2 * l + 1], lk.scharr_deriv(mine[l]))


EXTRA = np.array([[5, 5], [1236, 371], [0.4, 200.7], [1240.2, 3.3], [620.5, 375.9], [0, 0], [1240, 375]], np.float32)


def _assert_identical(A, B, pts):
    p1, st1, err1 = cv2.calcOpticalFlowPyrLK(A, B, pts.reshape(-1, 1, 2), None)
    p2, st2, err2, iters = lk.calc_optical_flow_pyr_lk(A, B, pts, return_iters=True)
    p1 = p1.reshape(-1, 2); st1 = st1.ravel(); err1 = err1.ravel()
    assert np.array_equal(st1, st2)
    ok = st2 == 1
    assert ok.sum() > 0.5 * len(pts)
    assert np.array_equal(p1[ok], p2[ok]), np.abs(p1 - p2)[ok].max()
    assert np.array_equal(err1[ok], err2[ok])
    assert iters.sum() > 0
    return p2, st2


@pytest.mark.parametrize("pair", ["temporal", "stereo"])
def test_lk_tracks_match_cv2(frames, pair):
    """1-channel tracks, grid step 9 (5,440 points) + border cases: bit-identical to cv2."""
    L0, L1, R0 = frames
    nxt = L1 if pair == "temporal" else R0
    pts = np.concatenate([glue.dense_keypoint_extractor(376, 1241, 9), EXTRA])
    p2, st2 = _assert_identical(L0, nxt, pts)
    # the round-1 variant (exact integer sums) stays within the tolerance but is not bit-identical
    sub = np.arange(0, len(pts), 5)
    p3, st3, _ = lk.calc_optical_flow_pyr_lk(L0, nxt, pts[sub], exact_sums=True)
    assert np.array_equal(st3, st2[sub])
    assert np.abs(p3 - p2[sub])[st3 == 1].max() <= 0.01


def test_lk_high_contrast_sums_beyond_2p24(frames):
    """Strong gradients: the float accumulators pass 2^24 and round at every step -- the regime in which
    the accumulation ORDER decides the bits (on the smooth synthetic frames most sums are still exact)."""
    L0, L1, _ = frames
    rng = np.random.default_rng(0)
    n = cv2.GaussianBlur(rng.integers(0, 2, L0.shape, dtype=np.uint8) * 255, (3, 3), 0.7)
    H0 = np.where(L0 > 110, n, 255 - n // 3).astype(np.uint8)
    M = np.float32([[1, 0, 1.3], [0, 1, 0.6]])
    H1 = cv2.warpAffine(H0, M, (H0.shape[1], H0.shape[0]), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    pts = glue.dense_keypoint_extractor(376, 1241, 20)
    _assert_identical(H0, H1, pts)


def _colorize(img):
    f = img.astype(np.float32)
    return np.stack([f, 255.0 - 0.8 * f, 255.0 * (f / 255.0) ** 0.7], -1).round().clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("kind", ["replicated_gray", "color"])
def test_lk_tracks_match_cv2_bgr(frames, kind):
    """3-channel images (what the reference passes, imread default): the restatement processes the
    channels as planes with one set of window sums; pinned against cv2 on the interleaved image."""
    L0, L1, _ = frames
    if kind == "replicated_gray":
        A, B = cv2.cvtColor(L0, cv2.COLOR_GRAY2BGR), cv2.cvtColor(L1, cv2.COLOR_GRAY2BGR)
    else:
        A, B = _colorize(L0), _colorize(L1)
    n, pyr = cv2.buildOpticalFlowPyramid(A, (21, 21), 3, withDerivatives=True)
    for c in range(3):
        mine = lk.build_pyramid(np.ascontiguousarray(A[:, :, c]), 3, 21)
        for l in range(4):
            assert np.array_equal(pyr[2 * l][:, :, c], mine[l])
            assert np.array_equal(pyr[2 * l + 1][:, :, 2 * c:2 * c + 2], lk.scharr_deriv(mine[l]))
    pts = np.concatenate([glue.dense_keypoint_extractor(376, 1241, 15), EXTRA])
    _assert_identical(A, B, pts)
