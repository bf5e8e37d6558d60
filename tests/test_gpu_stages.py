"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the
oracle (cv2 call-through + replay loops) on the same inputs, and against the committed golden
vectors.  Tolerances are the ones BASELINE.json's north_star states:
  keypoint sets / inlier index sets: bit-exact;  tracks: 0.01 px;  3-D points: 1e-4 relative;
  pose: 1e-4 rad, 1e-3 m.
"""
import cv2
import numpy as np
import pytest

from oracle import glue, lk as olk, replay, synth, cvrng
from gpu_common import golden, make_frontend

pytestmark = pytest.mark.gpu

TOL_PX = 0.01
TOL_REL3D = 1e-4
TOL_RAD = 1e-4
TOL_M = 1e-3


@pytest.fixture(scope="module")
def G():
    return golden()


@pytest.fixture(scope="module")
def fe():
    f = make_frontend()
    yield f
    f.close()


@pytest.fixture(scope="module")
def fe9():
    f = make_frontend(grid_step=9)
    yield f
    f.close()


def test_grid_keypoints_bit_exact(fe):
    img = np.zeros((376, 1241), np.uint8)
    for step in (30, 10, 9, 5, 2):
        a = fe.denseKeypointExtractor(img, step)
        b = glue.dense_keypoint_extractor(376, 1241, step)
        assert a.shape == b.shape and np.array_equal(a, b)
    assert len(fe.denseKeypointExtractor(np.zeros((40, 50), np.uint8), 30)) == 0


def test_pyramid_and_scharr_bit_exact(fe, G):
    L0 = G["L0"]
    n, pyr = cv2.buildOpticalFlowPyramid(L0, (21, 21), 3, withDerivatives=True)
    for l in range(4):
        lv, dv = fe.pyramid_level(L0, l)
        assert np.array_equal(lv, pyr[2 * l])
        assert np.array_equal(dv, pyr[2 * l + 1])
        assert int(lv.astype(np.int64).sum()) == int(G["pyr_sums"][l])
        assert int(np.abs(dv.astype(np.int64)).sum()) == int(G["deriv_abs_sums"][l])


def _padded_ref(img, levels, pad):
    n, pyr = cv2.buildOpticalFlowPyramid(img, (21, 21), levels, withDerivatives=True)
    out = []
    for l in range(n + 1):
        lv = cv2.copyMakeBorder(pyr[2 * l], pad, pad, pad, pad, cv2.BORDER_REFLECT_101)
        dv = cv2.copyMakeBorder(pyr[2 * l + 1], pad, pad, pad, pad, cv2.BORDER_CONSTANT, value=0)
        out.append((lv, dv))
    return out


def test_fused_pyramid_padded_borders_bit_exact(fe, G):
    """The one-launch TMA pyramid kernel writes every level, its REFLECT_101 border (21 px, what the
    LK window can reach) and the zero-bordered Scharr planes exactly as buildOpticalFlowPyramid does."""
    ref = _padded_ref(G["L0"], 3, 21)
    assert len(ref) == 4
    for l, (lv0, dv0) in enumerate(ref):
        lv, dv = fe.pyramid_padded(G["L0"], l, 21)
        assert np.array_equal(lv, lv0), "level %d" % l
        assert np.array_equal(dv, dv0), "deriv %d" % l


@pytest.mark.parametrize("size", [(333, 190), (64, 64), (257, 129), (1000, 64), (640, 480), (100, 70)])
def test_fused_pyramid_other_sizes(size):
    """Odd sizes, sizes below one CTA block, early stop of the level chain (next level <= winSize)."""
    w, h = size
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img = cv2.GaussianBlur(img, (5, 5), 0)
    f = make_frontend(width=w, height=h)
    try:
        ref = _padded_ref(img, 3, 21)
        for l, (lv0, dv0) in enumerate(ref):
            lv, dv = f.pyramid_padded(img, l, 21)
            assert np.array_equal(lv, lv0), "level %d of %dx%d" % (l, w, h)
            assert np.array_equal(dv, dv0), "deriv %d of %dx%d" % (l, w, h)
        with pytest.raises(Exception):
            f.pyramid_padded(img, len(ref), 21)       # no such level: the chain stopped where OpenCV's does
    finally:
        f.close()


@pytest.mark.parametrize("pair", ["temporal", "stereo"])
@pytest.mark.parametrize("step", [30, 9])
def test_lk_matches_opencv(fe, G, pair, step):
    L0 = G["L0"]
    nxt = G["L1"] if pair == "temporal" else G["R0"]
    pts = glue.dense_keypoint_extractor(376, 1241, step)
    extra = np.array([[5, 5], [1236, 371], [0.4, 200.7], [1240.2, 3.3], [620.5, 375.9]], np.float32)
    pts_x = np.concatenate([pts, extra])
    p, st, err = fe.calcOpticalFlowPyrLK(L0, nxt, pts_x)
    # (1) live cv2
    p0, st0, err0 = cv2.calcOpticalFlowPyrLK(L0, nxt, pts_x.reshape(-1, 1, 2), None)
    p0 = p0.reshape(-1, 2); st0 = st0.ravel(); err0 = err0.ravel()
    assert np.array_equal(st, st0)
    ok = st0 == 1
    # bit-identical: the kernel accumulates the window sums in OpenCV's float order (lk.cu)
    assert np.array_equal(p[ok], p0[ok]), np.abs(p - p0).max(1)[ok].max()
    assert np.array_equal(err[ok], err0[ok])
    # (2) committed golden vectors (grid part; generated with cv2 4.13.0 in the build container)
    okg = st[:len(pts)] == 1
    assert np.array_equal(st[:len(pts)], G[f"lk_{pair}_{step}_status"])
    assert np.array_equal(p[:len(pts)][okg], G[f"lk_{pair}_{step}_pts"][okg])
    assert np.array_equal(err[:len(pts)][okg], G[f"lk_{pair}_{step}_err"][okg])
    # (3) the scalar restatement (oracle/lk.py, OpenCV's float order)
    if step == 30:
        p1, st1, err1 = olk.calc_optical_flow_pyr_lk(L0, nxt, pts_x)
        assert np.array_equal(st, st1)
        assert np.array_equal(p[st1 == 1], p1[st1 == 1])
        assert np.array_equal(err[st1 == 1], err1[st1 == 1])


def test_lk_empty_and_out_of_image(fe, G):
    L0, L1 = G["L0"], G["L1"]
    p, st, err = fe.calcOpticalFlowPyrLK(L0, L1, np.zeros((0, 2), np.float32))
    assert len(p) == 0
    pts = np.array([[-50, -50], [2000, 100], [100, 900], [600, 180]], np.float32)
    p, st, err = fe.calcOpticalFlowPyrLK(L0, L1, pts)
    p0, st0, _ = cv2.calcOpticalFlowPyrLK(L0, L1, pts.reshape(-1, 1, 2), None)
    assert np.array_equal(st, st0.ravel())


def test_triangulate_bit_exact(fe):
    P1, P2 = glue.projection_matrices()
    rng = np.random.default_rng(0)
    n = 20000
    z = rng.uniform(3, 80, n)
    x1 = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], 1)
    x2 = x1.copy(); x2[:, 0] -= glue.FX * 0.54 / z
    x1 = (x1 + rng.normal(0, 0.1, (n, 2))).astype(np.float32)
    x2 = (x2 + rng.normal(0, 0.1, (n, 2))).astype(np.float32)
    a = fe.triangulatePoints(P1, P2, x1, x2)
    b = glue.triangulate(P1, P2, x1, x2)
    rel = np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-9)
    assert rel.max() <= TOL_REL3D
    assert np.array_equal(a, b)          # in fact bit-identical (same Jacobi SVD op order)
    assert len(fe.triangulatePoints(P1, P2, x1[:0], x2[:0])) == 0


def test_transform_points_bit_exact(fe):
    rng = np.random.default_rng(1)
    X = rng.uniform(-50, 50, (5000, 3)).astype(np.float32)
    pose = glue.camera_pose_from_pnp([0.01, -0.3, 0.02], [0.5, -0.1, 12.0])
    assert np.array_equal(fe.update3dtransformation(X, pose), glue.update_3d_transformation(X, pose))
    assert np.allclose(fe.pose_from_pnp([0.01, -0.3, 0.02], [0.5, -0.1, 12.0]), pose, rtol=0, atol=1e-15)


def _flow_case(n, seed, outlier_frac=0.2, grid=False):
    import test_oracle_ransac as t
    return t._flow_case(n, seed, outlier_frac, grid)


@pytest.mark.parametrize("n,seed,thr,grid", [(350, 0, 1.0, True), (350, 1, 3.0, True), (2000, 2, 1.0, False),
                                             (5000, 3, 3.0, False), (18000, 6, 1.0, False), (15, 5, 3.0, False)])
def test_fmat_ransac_mask_bit_exact(fe, n, seed, thr, grid):
    m1, m2 = _flow_case(n, seed, grid=grid)
    F0, mask0 = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, thr, 0.99)
    F, mask, ni = fe.findFundamentalMat(m1, m2, thr, 0.99)
    assert np.array_equal(mask, mask0.ravel())
    assert ni == int(mask0.sum())
    assert np.allclose(F, F0, rtol=1e-7, atol=1e-9)
    # replay: same minimal-sample index list on both sides
    r = replay.fmat_ransac(m1, m2, thr, 0.99)
    F2, mask2, _ = fe.findFundamentalMat(m1, m2, thr, 0.99, samples=r["samples"])
    assert np.array_equal(mask2, r["mask"])
    last = fe.last_fmat()
    assert last["best"] == r["best"]
    got = {(s, m): int(last["counts"][s, m]) for s in range(len(last["counts"])) for m in range(3) if last["counts"][s, m] >= 0}
    want = {(s, m): g for s, m, g in r["counts"]}
    assert got == want                      # every hypothesis has the same inlier count


def test_fmat_too_few_points(fe):
    m1, m2 = _flow_case(14, 0)
    from ros_stereo_slam_b200 import VoError
    with pytest.raises(VoError):
        fe.findFundamentalMat(m1[:6], m2[:6], 1.0)     # < 7 points: OpenCV returns nothing (7..14: test_small_point_sets_*)


@pytest.mark.parametrize("n,frac,iters,thr,conf", [(500, 0.1, 100, 1.0, 0.99), (5000, 0.3, 100, 1.0, 0.99),
                                                  (20000, 0.5, 100, 1.0, 0.99), (3000, 0.5, 400, 8.0, 0.98)])
def test_pnp_ransac_matches_opencv(fe, n, frac, iters, thr, conf):
    X, xy, _, _, _ = synth.pnp_stress_case(n, frac, 0.3, seed=3)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X.reshape(-1, 1, 3), xy.reshape(-1, 1, 2), glue.K, np.zeros((4, 1)),
                                             None, None, False, iters, thr, conf)
    r = fe.solvePnPRansac(X, xy, iters, thr, conf)
    assert r["ok"]
    assert np.array_equal(r["inliers"], inl.ravel())             # inlier index set bit-exact
    assert np.abs(r["rvec"] - rvec.ravel()).max() <= TOL_RAD
    assert np.abs(r["tvec"] - tvec.ravel()).max() <= TOL_M
    assert np.abs(r["rvec"] - rvec.ravel()).max() < 1e-7 and np.abs(r["tvec"] - tvec.ravel()).max() < 1e-6


def test_pnp_hypotheses_and_counts_replay(G):
    """Golden replay list: every hypothesis and every inlier count must match OpenCV's."""
    fe = make_frontend(ransac_exhaustive=1)
    X, xy = G["stress_X"], G["stress_xy"]
    S = G["stress_samples"]
    r = fe.solvePnPRansac(X, xy, 200, 1.0, 0.99, samples=S)
    last = fe.last_pnp()
    assert len(last["counts"]) == len(S)
    assert np.array_equal(last["counts"], G["stress_counts"])
    assert np.abs(last["models"] - G["stress_hyp"]).max() < 1e-9
    assert last["best"] == int(G["stress_best"]) and last["n_iters"] == int(G["stress_niters"])
    assert np.array_equal(r["inliers"], G["stress_inliers"])
    assert np.abs(r["rvec"] - G["stress_rvec"]).max() < 1e-7 and np.abs(r["tvec"] - G["stress_tvec"]).max() < 1e-6
    # exhaustive and early-exit evaluation give the same answer
    fe2 = make_frontend(ransac_exhaustive=0)
    r2 = fe2.solvePnPRansac(X, xy, 200, 1.0, 0.99)
    assert np.array_equal(r2["inliers"], r["inliers"]) and np.array_equal(r2["rvec"], r["rvec"])
    fe.close(); fe2.close()


def test_pnp_default_sample_stream(fe):
    X, xy, _, _, _ = synth.pnp_stress_case(500, 0.1, 0.3, seed=3)
    fe.solvePnPRansac(X, xy, 100, 1.0, 0.99)
    # first sample drawn for N=500
    assert list(cvrng.sample_list(500, 5, 1)[0]) == [105, 4, 440, 173, 331]


@pytest.mark.parametrize("step", [30, 9])
def test_stage_boundaries_match_reference_glue(G, step):
    fe = make_frontend(grid_step=step)
    L0, R0, L1, R1 = G["L0"], G["R0"], G["L1"], G["R1"]
    # stereoTriangulate
    xyz, ref2d = fe.stereoTriangulate(L0, R0)
    xyz0, ref0 = glue.stereo_triangulate(L0, R0, step)
    assert np.array_equal(ref2d, ref0)                       # sampled keypoint set bit-exact
    assert np.array_equal(ref2d, G[f"stereo_{step}_ref2d"])
    rel = np.abs(xyz - xyz0).max(1) / np.abs(xyz0).max(1)
    assert rel.max() <= TOL_REL3D
    # PerspectiveNpointEstimation on the reference's own inputs
    res = fe.PerspectiveNpointEstimation(L0, L1, ref0, xyz0)
    ref = glue.perspective_n_point_estimation(L0, L1, ref0, xyz0, iters=100)
    assert res["trk2d"].shape == ref["trk2d"].shape
    assert np.abs(res["trk2d"] - ref["trk2d"]).max() <= TOL_PX
    assert np.array_equal(res["trk3d"], ref["trk3d"])
    assert np.array_equal(res["ref2d_inl"], ref["ref2d_inl"])
    assert res["attempt"] == ref["attempt"] and res["shutdown"] == ref["shutdown"]
    assert np.array_equal(res["inliers"], ref["inliers"])
    assert np.abs(res["rvec"] - ref["rvec"]).max() <= TOL_RAD
    assert np.abs(res["tvec"] - ref["tvec"]).max() <= TOL_M
    assert np.array_equal(res["inliers"], G[f"pnp_{step}_inliers"])
    # insertKeyFrames
    pose = glue.camera_pose_from_pnp(ref["rvec"], ref["tvec"])
    w3, w2, cam = fe.insertKeyFrames(L1, R1, pose)
    w3r, w2r, camr = glue.insert_key_frames(L1, R1, pose, step)
    assert np.array_equal(w2, w2r)
    assert (np.abs(w3 - w3r).max(1) / np.abs(w3r).max(1)).max() <= TOL_REL3D
    fe.close()


def test_anms_set_bit_exact(fe):
    rng = np.random.default_rng(5)
    n = 3000
    xy = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], 1).astype(np.float32)
    resp = rng.uniform(0, 1, n).astype(np.float32)
    for keep in (100, 1000, 2999):
        a = fe.adaptiveNonMaximalSuppresion(xy, resp, keep)
        b = glue.anms(xy, resp, keep)
        assert np.array_equal(a, b)
    # grid keypoints have response 0 -> everything is kept
    g = glue.dense_keypoint_extractor(376, 1241, 30)
    a = fe.adaptiveNonMaximalSuppresion(g, np.zeros(len(g), np.float32), 100)
    assert np.array_equal(np.sort(a), np.arange(len(g)))
    # fewer than numToKeep: returned unchanged
    assert np.array_equal(fe.adaptiveNonMaximalSuppresion(xy[:50], resp[:50], 100), np.arange(50))


def test_sequence_driver_matches_reference_loop():
    """Device-resident sequence driver vs the reference frame loop restated over cv2."""
    sc = synth.Scene(1)
    n = 4
    Ls = [sc.render(i, "L") for i in range(n)]
    Rs = [sc.render(i, "R") for i in range(n)]
    for kf in (200, 10 ** 9):
        fe = make_frontend(kf_min_inliers=min(kf, 2 ** 31 - 1))
        ref = glue.run_sequence(Ls, Rs, step=30, pnp_iters=100, kf_min_inliers=kf)
        n0 = fe.seq_init(Ls[0], Rs[0])
        for i in range(1, n):
            res, code = fe.seq_track(Ls[i], Rs[i])
            want = ref[i - 1]
            assert res.n_lk_in == want["n_lk_in"] and res.n_tracked == want["n_tracked"]
            assert res.n_inliers == want["n_inliers"] and bool(res.keyframe) == want["keyframe"]
            assert np.abs(np.array(res.rvec) - want["rvec"]).max() <= TOL_RAD
            assert np.abs(np.array(res.tvec) - want["tvec"]).max() <= TOL_M
        fe.close()


# ---------------------------------------------------------------------------- 3-channel (BGR) input, SURVEY 8(f)-1
def _colorize(img, k):
    """A BGR image whose channels really differ (gain / inversion / gamma of the gray frame)."""
    f = img.astype(np.float32)
    b = f
    g = 255.0 - 0.8 * f
    r = 255.0 * (f / 255.0) ** (0.7 + 0.1 * k)
    return np.stack([b, g, r], -1).round().clip(0, 255).astype(np.uint8)


@pytest.fixture(scope="module")
def fe3():
    f = make_frontend(channels=3)
    yield f
    f.close()


@pytest.mark.parametrize("kind", ["replicated_gray", "color"])
def test_bgr_pyramid_and_scharr_bit_exact(fe3, G, kind):
    """channels=3: levels, REFLECT_101 borders and the 6-channel Scharr planes equal
    cv::buildOpticalFlowPyramid on the interleaved BGR image (imread's output, reference
    src/keyFrameManagement.cpp:52)."""
    img = cv2.cvtColor(G["L0"], cv2.COLOR_GRAY2BGR) if kind == "replicated_gray" else _colorize(G["L0"], 0)
    ref = _padded_ref(img, 3, 21)
    assert len(ref) == 4
    for l, (lv0, dv0) in enumerate(ref):
        lv, dv = fe3.pyramid_padded(img, l, 21)
        assert lv.shape == lv0.shape and dv.shape == dv0.shape
        assert np.array_equal(lv, lv0), "level %d" % l
        assert np.array_equal(dv, dv0), "deriv %d" % l


@pytest.mark.parametrize("kind", ["replicated_gray", "color"])
@pytest.mark.parametrize("pair", ["temporal", "stereo"])
def test_bgr_lk_matches_opencv(fe3, G, kind, pair):
    """3-channel LK against cv2.calcOpticalFlowPyrLK on the same BGR images: status, positions and err bit-identical."""
    a, b = G["L0"], (G["L1"] if pair == "temporal" else G["R0"])
    if kind == "replicated_gray":
        A, B = cv2.cvtColor(a, cv2.COLOR_GRAY2BGR), cv2.cvtColor(b, cv2.COLOR_GRAY2BGR)
    else:
        A, B = _colorize(a, 0), _colorize(b, 0)
    pts = glue.dense_keypoint_extractor(376, 1241, 9)
    extra = np.array([[5, 5], [1236, 371], [0.4, 200.7], [1240.2, 3.3], [620.5, 375.9]], np.float32)
    pts = np.concatenate([pts, extra])
    p, st, err = fe3.calcOpticalFlowPyrLK(A, B, pts)
    p0, st0, err0 = cv2.calcOpticalFlowPyrLK(A, B, pts.reshape(-1, 1, 2), None)
    p0 = p0.reshape(-1, 2); st0 = st0.ravel(); err0 = err0.ravel()
    assert np.array_equal(st, st0)
    ok = st0 == 1
    # bit-identical: three warps per keypoint accumulate the interleaved window sums in OpenCV's float order (lk.cu)
    assert np.array_equal(p[ok], p0[ok]), np.abs(p - p0).max(1)[ok].max()
    assert np.array_equal(err[ok], err0[ok])
    # and so is the multi-channel scalar restatement (oracle/lk.py, pinned against cv2 in tests/test_oracle_lk.py)
    sub = np.r_[0:len(pts):7, len(pts) - 5:len(pts)]
    p1, st1, err1 = olk.calc_optical_flow_pyr_lk(A, B, pts[sub])
    assert np.array_equal(st[sub], st1)
    assert np.array_equal(p[sub][st1 == 1], p1[st1 == 1])
    assert np.array_equal(err[sub][st1 == 1], err1[st1 == 1])
    # the 3-channel result is NOT the 1-channel one: the path really uses all channels
    if kind == "color":
        f1 = make_frontend()
        p1, st1, _ = f1.calcOpticalFlowPyrLK(a, b, pts)
        f1.close()
        both = (st1 == 1) & ok
        assert np.abs(p1[both] - p[both]).max() > 1e-3


def test_bgr_to_gray_bit_exact(fe, G):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (376, 1241, 3), dtype=np.uint8)
    assert np.array_equal(fe.cvtColorBGR2GRAY(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    col = _colorize(G["L0"], 1)
    assert np.array_equal(fe.cvtColorBGR2GRAY(col), glue.bgr_to_gray(col))
    # replicated gray comes back unchanged (3735 + 19235 + 9798 = 2^15)
    assert np.array_equal(fe.cvtColorBGR2GRAY(cv2.cvtColor(G["L0"], cv2.COLOR_GRAY2BGR)), G["L0"])


def test_bgr_stage_entry_points_and_sequence(G):
    """stereoTriangulate / PerspectiveNpointEstimation / sequence driver on BGR frames vs the oracle glue
    (the reference's own data flow: imread -> LK on 3 channels)."""
    L0, R0, L1 = [cv2.cvtColor(G[k], cv2.COLOR_GRAY2BGR) for k in ("L0", "R0", "L1")]
    fe = make_frontend(channels=3)
    xyz, ref2d = fe.stereoTriangulate(L0, R0)
    xyz0, ref0 = glue.stereo_triangulate(L0, R0, 30)
    assert np.array_equal(ref2d, ref0)
    assert (np.abs(xyz - xyz0).max(1) / np.abs(xyz0).max(1)).max() <= TOL_REL3D
    res = fe.PerspectiveNpointEstimation(L0, L1, ref0, xyz0)
    ref = glue.perspective_n_point_estimation(L0, L1, ref0, xyz0, iters=100)
    assert np.array_equal(res["inliers"], ref["inliers"])
    assert np.abs(res["rvec"] - ref["rvec"]).max() <= TOL_RAD and np.abs(res["tvec"] - ref["tvec"]).max() <= TOL_M
    fe.close()
    sc = synth.Scene(1)
    Ls = [cv2.cvtColor(sc.render(i, "L"), cv2.COLOR_GRAY2BGR) for i in range(4)]
    Rs = [cv2.cvtColor(sc.render(i, "R"), cv2.COLOR_GRAY2BGR) for i in range(4)]
    fe = make_frontend(channels=3, kf_min_inliers=2 ** 31 - 1)
    want = glue.run_sequence(Ls, Rs, step=30, pnp_iters=100, kf_min_inliers=10 ** 9)
    fe.seq_init(Ls[0], Rs[0])
    for i in range(1, 4):
        res, code = fe.seq_track(Ls[i], Rs[i])
        w = want[i - 1]
        assert res.n_lk_in == w["n_lk_in"] and res.n_tracked == w["n_tracked"] and res.n_inliers == w["n_inliers"]
        assert np.abs(np.array(res.rvec) - w["rvec"]).max() <= TOL_RAD
        assert np.abs(np.array(res.tvec) - w["tvec"]).max() <= TOL_M
    fe.close()


# ---------------------------------------------------------------------------- SORcloud, SURVEY 8(f)-3
def test_sor_cloud_matches_restatement(fe, G):
    """visualSLAM::SORcloud on real keyframe clouds (stereoTriangulate at grid step 9 and 5) and on a synthetic
    one with far outliers, duplicates and a -z > 500 point: kept index set identical to oracle/sor.py, per-point
    mean neighbour distances bit-identical."""
    from oracle import sor as osor
    clouds = []
    for step in (9, 5):
        f = make_frontend(grid_step=step)
        xyz, _ = f.stereoTriangulate(G["L0"], G["R0"])
        f.close()
        clouds.append(xyz)
    rng = np.random.default_rng(11)
    syn = np.c_[rng.uniform(-10, 10, 4000), rng.normal(1.65, 0.05, 4000), rng.uniform(4, 60, 4000)].astype(np.float32)
    syn[100:120] = syn[99]                       # duplicates: several zero distances
    syn[7] = (1, 2, -700)                        # never enters the cloud
    syn[3000:3050] += rng.uniform(50, 90, (50, 3)).astype(np.float32)
    clouds.append(syn)
    # point order must not matter for the per-point distances
    clouds.append(clouds[0][rng.permutation(len(clouds[0]))])
    for xyz in clouds:
        assert len(xyz) > 1000
        pts, _c, idx, md = fe.SORcloud(xyz, None, 200, 0.01, return_distances=True)
        keep0, dist0, thr0 = osor.sor_cloud(xyz, 200, 0.01, return_all=True)
        assert np.array_equal(md, dist0)
        assert np.array_equal(idx, keep0)
        assert np.array_equal(pts, xyz[keep0])
        assert 0 < len(idx) < len(xyz)
    # colours follow the points; fewer points than meanK + 1 keeps everything that enters the cloud
    small = clouds[0][:150]
    col = np.arange(450, dtype=np.float32).reshape(150, 3)
    pts, c2 = fe.SORcloud(small, col)
    assert np.array_equal(pts, small) and np.array_equal(c2, col)
    pts, _ = fe.SORcloud(np.zeros((0, 3), np.float32))
    assert len(pts) == 0
    # other meanK values
    for k in (600, 20):
        _p, _c, idx, md = fe.SORcloud(clouds[0], None, k, 0.5, return_distances=True)
        keep0, dist0, _t = osor.sor_cloud(clouds[0], k, 0.5, return_all=True)
        assert np.array_equal(md, dist0) and np.array_equal(idx, keep0)


def test_sequence_prefetch_is_a_pure_transfer_overlap():
    """vo_seq_prefetch only moves the H2D copy of the next frame under the current frame's processing:
    every field of every frame result is identical with and without it."""
    sc = synth.Scene(4)
    n = 4
    Ls = [np.ascontiguousarray(sc.render(i, "L")) for i in range(n)]
    Rs = [np.ascontiguousarray(sc.render(i, "R")) for i in range(n)]

    def fields(res):
        return (res.n_lk_in, res.n_tracked, res.n_inliers, res.attempt_used, res.keyframe, res.n_kf_points,
                res.n_lk_in_stereo, tuple(res.rvec), tuple(res.tvec), tuple(res.pose3x4))

    for kf in (200, 2 ** 31 - 1):
        out = []
        for use_prefetch in (False, True):
            fe = make_frontend(kf_min_inliers=kf)
            fe.seq_init(Ls[0], Rs[0])
            rows = []
            if use_prefetch:
                fe.seq_prefetch(Ls[1], Rs[1])
            for i in range(1, n):
                if use_prefetch and i + 1 < n:
                    fe.seq_prefetch(Ls[i + 1], Rs[i + 1])
                res, code = fe.seq_track(Ls[i], Rs[i])
                assert code == 0
                rows.append(fields(res))
            out.append((rows, fe.seq_reference()))
            fe.close()
        assert out[0][0] == out[1][0]
        assert np.array_equal(out[0][1][0], out[1][1][0]) and np.array_equal(out[0][1][1], out[1][1][1])


# ---------------------------------------------------------------------------- full-size cases
def test_full_size_config2_lk_and_stages(G):
    """BASELINE config 2 sizes: grid step 5 (18,278 keypoints), 1024 PnP hypotheses."""
    fe = make_frontend(grid_step=5, pnp_iters=1024, ransac_exhaustive=1)
    L0, R0, L1 = G["L0"], G["R0"], G["L1"]
    pts = glue.dense_keypoint_extractor(376, 1241, 5)
    assert len(pts) == 18278
    p, st, err = fe.calcOpticalFlowPyrLK(L0, R0, pts)
    p0, st0, _ = cv2.calcOpticalFlowPyrLK(L0, R0, pts.reshape(-1, 1, 2), None)
    assert np.array_equal(st, st0.ravel())
    assert np.array_equal(p[st == 1], p0.reshape(-1, 2)[st == 1])        # every track bit-identical to cv2
    xyz, ref2d = fe.stereoTriangulate(L0, R0)
    xyz0, ref0 = glue.stereo_triangulate(L0, R0, 5)
    assert np.array_equal(ref2d, ref0)
    assert (np.abs(xyz - xyz0).max(1) / np.abs(xyz0).max(1)).max() <= TOL_REL3D
    res = fe.PerspectiveNpointEstimation(L0, L1, ref0, xyz0)
    ref = glue.perspective_n_point_estimation(L0, L1, ref0, xyz0, iters=1024)
    # identical tracks into identical RANSAC loops: the chained sets are equalities
    assert np.array_equal(res["trk2d"], ref["trk2d"])
    assert np.array_equal(res["inliers"], ref["inliers"])
    assert np.abs(res["rvec"] - ref["rvec"]).max() <= TOL_RAD
    assert np.abs(res["tvec"] - ref["tvec"]).max() <= TOL_M
    fe.close()


def test_density_stress_config3():
    """BASELINE config 3: step-2 candidates (115,134) -> ANMS(80000) -> LK on the kept set."""
    fe = make_frontend(max_points=131072)
    sc = synth.Scene(2)
    L0, L1 = sc.render(0, "L"), sc.render(1, "L")
    cand = glue.dense_keypoint_extractor(376, 1241, 2)
    assert len(cand) == 115134
    # response = a cheap texture measure (local gradient energy) so that ANMS has something to rank
    gx = cv2.Sobel(L0, cv2.CV_32F, 1, 0, ksize=3); gy = cv2.Sobel(L0, cv2.CV_32F, 0, 1, ksize=3)
    resp = cv2.boxFilter(gx * gx + gy * gy, -1, (7, 7))[cand[:, 1].astype(int), cand[:, 0].astype(int)].astype(np.float32)
    keep = fe.adaptiveNonMaximalSuppresion(cand, resp, 80000)
    assert len(keep) >= 80001 and len(np.unique(keep)) == len(keep)
    # property: kept set is closed under the ANMS rule -- checked against the oracle on a subsample
    sub = np.arange(0, len(cand), 23)
    assert np.array_equal(fe.adaptiveNonMaximalSuppresion(cand[sub], resp[sub], 3000), glue.anms(cand[sub], resp[sub], 3000))
    pts = cand[keep]
    p, st, err = fe.calcOpticalFlowPyrLK(L0, L1, pts)
    p0, st0, _ = cv2.calcOpticalFlowPyrLK(L0, L1, pts.reshape(-1, 1, 2), None)
    assert np.array_equal(st, st0.ravel())
    assert np.array_equal(p[st == 1], p0.reshape(-1, 2)[st == 1])        # ~90k tracks, bit-identical to cv2
    fe.close()


def test_pnp_stress_config4(fe):
    """BASELINE config 4: N = 20,000, 50 % outliers, 4096 iterations, conf 0.99 (OpenCV's adaptive rule stops
    early; parity is defined on those semantics) and the exhaustive evaluation of all 4096 hypotheses."""
    X, xy, _, _, _ = synth.pnp_stress_case(20000, 0.5, 0.3, seed=3)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X.reshape(-1, 1, 3), xy.reshape(-1, 1, 2), glue.K, np.zeros((4, 1)),
                                             None, None, False, 4096, 1.0, 0.99)
    for ex in (0, 1):
        f = make_frontend(ransac_exhaustive=ex)
        r = f.solvePnPRansac(X, xy, 4096, 1.0, 0.99)
        assert np.array_equal(r["inliers"], inl.ravel())
        assert np.abs(r["rvec"] - rvec.ravel()).max() <= TOL_RAD and np.abs(r["tvec"] - tvec.ravel()).max() <= TOL_M
        if ex:
            assert len(f.last_pnp()["counts"]) == 4096
        f.close()


def test_edge_cases(fe, G):
    from ros_stereo_slam_b200 import VoError
    L0 = G["L0"]
    # empty inputs
    assert len(fe.update3dtransformation(np.zeros((0, 3), np.float32), np.eye(3, 4))) == 0
    a, b = fe.denseLKtracking(L0, G["L1"], np.zeros((0, 2), np.float32))
    assert len(a) == 0 and len(b) == 0
    # fewer than four correspondences: OpenCV asserts (5 and 4 points take its direct EPnP / P3P solve, tested below)
    with pytest.raises(VoError):
        fe.solvePnPRansac(np.zeros((3, 3), np.float32), np.zeros((3, 2), np.float32))
    # all-outlier PnP: no model
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (50, 3)).astype(np.float32); X[:, 2] += 20
    xy = rng.uniform(0, 1000, (50, 2)).astype(np.float32)
    ok0, _, _, inl0 = cv2.solvePnPRansac(X.reshape(-1, 1, 3), xy.reshape(-1, 1, 2), glue.K, np.zeros((4, 1)),
                                         None, None, False, 100, 1.0, 0.99)
    r = fe.solvePnPRansac(X, xy, 100, 1.0, 0.99)
    n0 = 0 if inl0 is None else len(inl0)
    assert r["ok"] == bool(ok0) and len(r["inliers"]) == n0
    # textureless image: LK rejects everything (minEig) exactly like OpenCV
    flat = np.full((376, 1241), 100, np.uint8)
    pts = glue.dense_keypoint_extractor(376, 1241, 30)
    p, st, err = fe.calcOpticalFlowPyrLK(flat, flat, pts)
    p0, st0, _ = cv2.calcOpticalFlowPyrLK(flat, flat, pts.reshape(-1, 1, 2), None)
    assert np.array_equal(st, st0.ravel()) and st.sum() == 0
    # strided (non-contiguous) image views are accepted
    big = np.zeros((376, 1300), np.uint8); big[:, :1241] = L0
    lv, dv = fe.pyramid_level(big[:, :1241], 1)
    assert np.array_equal(lv, fe.pyramid_level(L0, 1)[0])
    # self-check entry point
    assert fe.lib.vo_self_check(fe.h) == 0


def test_gpu_renderer_matches_oracle_renderer(fe):
    """The bench renders its sequences with the harness kernel (vo_synth_render_dev) and the reference arm with
    oracle/synth.py: same scene description, same trajectory.  The two evaluate the texture in float vs double,
    so grey levels may differ by one and isolated silhouette pixels may pick the other surface."""
    stats = []
    for seed, frame, eye in ((0, 0, 0), (0, 7, 1), (3, 2, 0)):
        g = fe.synth_render(seed, frame, eye)
        o = synth.Scene(seed).render(frame, "L" if eye == 0 else "R")
        assert g.shape == o.shape == (376, 1241)
        d = np.abs(g.astype(np.int32) - o.astype(np.int32))
        stats.append((float(np.mean(d == 0)), float(np.mean(d <= 1)), int(d.max())))
    print("renderer agreement (identical, within 1 level, max):", stats)
    for ident, near, _mx in stats:
        assert ident >= 0.95 and near >= 0.999, stats


def test_textureless_first_frame_on_a_fresh_context():
    """with fewer LK survivors than a minimal sample the device-side sampler leaves its output to the
    solve kernels of the fused chain -- on a FRESH context (nothing ever written to the sample buffer) that must
    not turn into a wild gather.  A constant image has no trackable point at all."""
    flat = np.full((376, 1241), 90, np.uint8)
    for _ in range(2):
        f = make_frontend()
        xyz, ref2d = f.stereoTriangulate(flat, flat)
        assert len(xyz) == 0 and len(ref2d) == 0
        assert f.seq_init(flat, flat) == 0
        f.close()
    f = make_frontend()
    pts = glue.dense_keypoint_extractor(376, 1241, 30)
    res = f.PerspectiveNpointEstimation(flat, flat, pts, np.ones((len(pts), 3), np.float32))
    assert res["shutdown"] and len(res["trk2d"]) == 0 and len(res["inliers"]) == 0      # the reference's SHUTDOWN_FLAG
    # the context is still healthy afterwards
    g = golden()
    xyz, ref2d = f.stereoTriangulate(g["L0"], g["R0"])
    assert len(xyz) > 300
    f.close()


def test_create_rejects_capacities_the_fused_chains_would_overrun():
    with pytest.raises(Exception):
        make_frontend(max_hypotheses=32)                       # below the 96-sample first chunk of the temporal F-RANSAC
    with pytest.raises(Exception):
        make_frontend(max_hypotheses=512, pnp_iters=1024, ransac_exhaustive=1)
    with pytest.raises(Exception):
        make_frontend(lk_max_level=4)                          # would silently be clamped to 4 levels
    f = make_frontend(max_hypotheses=96, max_points=600000)     # more compaction tiles than the old fixed table held
    f.close()


def test_stage_call_ends_the_sequence(G):
    """The stage entry points build their pyramids in the sequence driver's slots: a sequence must be re-initialised."""
    f = make_frontend()
    assert f.seq_init(G["L0"], G["R0"]) > 300
    f.calcOpticalFlowPyrLK(G["L0"], G["L1"], glue.dense_keypoint_extractor(376, 1241, 30))
    with pytest.raises(Exception):
        f.seq_track(G["L1"], G["R0"])
    assert f.seq_init(G["L0"], G["R0"]) > 300
    res, code = f.seq_track(G["L1"], G["R0"])
    assert res.n_inliers > 100
    f.close()


def test_wrapper_rejects_images_of_another_geometry(fe, G):
    with pytest.raises(ValueError):
        fe.calcOpticalFlowPyrLK(G["L0"][:300], G["L1"][:300], np.zeros((4, 2), np.float32))
    with pytest.raises(ValueError):
        fe.stereoTriangulate(cv2.cvtColor(G["L0"], cv2.COLOR_GRAY2BGR), cv2.cvtColor(G["R0"], cv2.COLOR_GRAY2BGR))
    # a row-strided left view next to a contiguous right image: both are brought to one stride
    wide = np.zeros((376, 1300), np.uint8)
    wide[:, :1241] = G["L0"]
    p, st, _ = fe.calcOpticalFlowPyrLK(wide[:, :1241], G["L1"], glue.dense_keypoint_extractor(376, 1241, 30))
    p0, st0, _ = fe.calcOpticalFlowPyrLK(G["L0"], G["L1"], glue.dense_keypoint_extractor(376, 1241, 30))
    assert np.array_equal(st, st0) and np.array_equal(p, p0)


@pytest.mark.parametrize("channels", [1, 3])
def test_lookahead_changes_the_schedule_not_the_numbers(channels):
    """Keyframe on every frame + an announced next frame: its temporal LK runs one call early on the look-ahead chain
    (all stereo-LK survivors, inliers gathered afterwards).  Every per-frame result and the final reference set must be
    identical to the run without announcements, for device-resident (vo_seq_announce) and host (vo_seq_prefetch)
    frames, and when the caller then passes a DIFFERENT frame than the one it announced."""
    import ctypes as C
    from ros_stereo_slam_b200 import _lib
    sc = synth.Scene(4)
    n = 7
    Ls = [sc.render(i, "L") for i in range(n)]
    Rs = [sc.render(i, "R") for i in range(n)]
    if channels == 3:
        Ls = [cv2.cvtColor(x, cv2.COLOR_GRAY2BGR) for x in Ls]
        Rs = [cv2.cvtColor(x, cv2.COLOR_GRAY2BGR) for x in Rs]

    def fields(res):
        return (res.n_lk_in, res.n_tracked, res.n_inliers, res.attempt_used, res.keyframe, res.n_kf_points,
                res.n_lk_in_stereo, tuple(res.rvec), tuple(res.tvec), tuple(res.pose3x4))

    order = [1, 2, 3, 5, 4, 6]          # frame 5 arrives where 4 was announced: that look-ahead is discarded

    def run(mode):
        import os
        # the look-ahead is opt-in and lives in the host-driven chains; the switches are read at vo_create
        os.environ["VO_B200_SEQ_HOST"] = "1"
        os.environ["VO_B200_LOOKAHEAD"] = "1"
        try:
            fe = make_frontend(kf_min_inliers=2 ** 31 - 1, channels=channels, grid_step=9)
        finally:
            del os.environ["VO_B200_SEQ_HOST"], os.environ["VO_B200_LOOKAHEAD"]
        nb = Ls[0].nbytes
        d = C.c_void_p()
        if mode == "device":
            _lib.check(fe.lib.vo_alloc_dev(fe.h, C.byref(d), C.c_uint64(2 * n * nb)))
            for i in range(n):
                for e, img in enumerate((Ls[i], Rs[i])):
                    _lib.check(fe.lib.vo_memcpy_h2d(fe.h, C.c_void_p(d.value + (2 * i + e) * nb), img.ctypes.data_as(C.c_void_p), C.c_uint64(nb)))
        dp = lambda i, e: d.value + (2 * i + e) * nb
        stride = Ls[0].strides[0]
        if mode == "device":
            fe.seq_init(dp(0, 0), dp(0, 1), is_device=True, stride=stride)
        else:
            fe.seq_init(Ls[0], Rs[0])
        rows = []
        for j, i in enumerate(order):
            announced = i + 1 if j + 1 < len(order) and i + 1 < n else None     # always the NEXT INDEX, not the next in `order`
            if announced is not None and mode == "device":
                fe.seq_announce(dp(announced, 0), dp(announced, 1), stride)
            elif announced is not None and mode == "host":
                fe.seq_prefetch(Ls[announced], Rs[announced])
            if mode == "device":
                res, code = fe.seq_track(dp(i, 0), dp(i, 1), is_device=True, stride=stride)
            else:
                res, code = fe.seq_track(Ls[i], Rs[i])
            assert code == 0
            rows.append(fields(res))
        ref = fe.seq_reference()
        launches = fe.launch_count()
        if mode == "device":
            fe.lib.vo_free_dev(fe.h, d)
        fe.close()
        return rows, ref, launches

    plain = run("plain")
    for mode in ("device", "host"):
        got = run(mode)
        assert got[0] == plain[0], mode
        assert np.array_equal(got[1][0], plain[1][0]) and np.array_equal(got[1][1], plain[1][1]), mode


def test_pyramid_ahead_changes_the_schedule_not_the_numbers():
    """Default (fused) chains + an announced next frame: its LEFT pyramid is built one call early on the third stream.
    Every per-frame result and the final reference set must equal the run without announcements, for device-resident
    (vo_seq_announce) and host (vo_seq_prefetch) frames, and when the caller passes another frame than it announced."""
    import ctypes as C
    from ros_stereo_slam_b200 import _lib
    sc = synth.Scene(4)
    n = 7
    Ls = [sc.render(i, "L") for i in range(n)]
    Rs = [sc.render(i, "R") for i in range(n)]

    def fields(res):
        return (res.n_lk_in, res.n_tracked, res.n_inliers, res.attempt_used, res.keyframe, res.n_kf_points,
                res.n_lk_in_stereo, tuple(res.rvec), tuple(res.tvec), tuple(res.pose3x4))

    order = [1, 2, 3, 5, 4, 6]          # frame 5 arrives where 4 was announced: that pyramid is discarded

    def run(mode):
        fe = make_frontend(kf_min_inliers=2 ** 31 - 1, grid_step=9)
        nb = Ls[0].nbytes
        d = C.c_void_p()
        if mode == "device":
            _lib.check(fe.lib.vo_alloc_dev(fe.h, C.byref(d), C.c_uint64(2 * n * nb)))
            for i in range(n):
                for e, img in enumerate((Ls[i], Rs[i])):
                    _lib.check(fe.lib.vo_memcpy_h2d(fe.h, C.c_void_p(d.value + (2 * i + e) * nb), img.ctypes.data_as(C.c_void_p), C.c_uint64(nb)))
        dp = lambda i, e: d.value + (2 * i + e) * nb
        stride = Ls[0].strides[0]
        if mode == "device":
            fe.seq_init(dp(0, 0), dp(0, 1), is_device=True, stride=stride)
        else:
            fe.seq_init(Ls[0], Rs[0])
        rows = []
        for j, i in enumerate(order):
            announced = i + 1 if j + 1 < len(order) and i + 1 < n else None
            if announced is not None and mode == "device":
                fe.seq_announce(dp(announced, 0), dp(announced, 1), stride)
            elif announced is not None and mode == "host":
                fe.seq_prefetch(Ls[announced], Rs[announced])
            if mode == "device":
                res, code = fe.seq_track(dp(i, 0), dp(i, 1), is_device=True, stride=stride)
            else:
                res, code = fe.seq_track(Ls[i], Rs[i])
            assert code == 0
            rows.append(fields(res))
        ref = fe.seq_reference()
        # a stage call while a pyramid built ahead may still be in flight must not disturb it (and ends the sequence)
        fe.denseKeypointExtractor(Ls[0], 30)
        if mode == "device":
            fe.lib.vo_free_dev(fe.h, d)
        fe.close()
        return rows, ref

    plain = run("plain")
    for mode in ("device", "host"):
        got = run(mode)
        assert got[0] == plain[0], mode
        assert np.array_equal(got[1][0], plain[1][0]) and np.array_equal(got[1][1], plain[1][1]), mode


def test_sm_partition_changes_the_schedule_not_the_numbers():
    """VO_B200_ISLAND=8 (CUDA green contexts): the LK launches of both chains run on 140 SMs, the tracking chain's
    latency-bound F-RANSAC kernels on an 8-SM island no LK launch can occupy.  Every frame result must equal the
    default schedule's."""
    import os
    sc = synth.Scene(4)
    n = 5
    Ls = [np.ascontiguousarray(sc.render(i, "L")) for i in range(n)]
    Rs = [np.ascontiguousarray(sc.render(i, "R")) for i in range(n)]

    def fields(res):
        return (res.n_lk_in, res.n_tracked, res.n_inliers, res.attempt_used, res.keyframe, res.n_kf_points,
                res.n_lk_in_stereo, tuple(res.rvec), tuple(res.tvec), tuple(res.pose3x4))

    out = []
    for island in (None, "8"):
        if island:
            os.environ["VO_B200_ISLAND"] = island
        try:
            fe = make_frontend(kf_min_inliers=2 ** 31 - 1, grid_step=9)
        finally:
            os.environ.pop("VO_B200_ISLAND", None)
        fe.seq_init(Ls[0], Rs[0])
        rows = []
        for i in range(1, n):
            res, code = fe.seq_track(Ls[i], Rs[i])
            assert code == 0
            rows.append(fields(res))
        out.append((rows, fe.seq_reference()))
        fe.close()
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1][0], out[1][1][0]) and np.array_equal(out[0][1][1], out[1][1][1])


def test_small_point_sets_take_opencv_paths(fe, G):
    """SURVEY 8 a-5 / a-7 small-N behaviour (missing items 1-2): findFundamentalMat with 7 and 8..14 points,
    solvePnPRansac with 5 and 4 points, against the live cv2."""
    L0, L1 = G["L0"], G["L1"]
    pts = glue.dense_keypoint_extractor(376, 1241, 30)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(L0, L1, pts.reshape(-1, 1, 2), None)
    ok = st.ravel() == 1
    a, b = pts[ok], p1.reshape(-1, 2)[ok]
    rng = np.random.default_rng(0)
    same = total = 0
    for n in range(8, 15):
        for trial in range(6):
            sel = rng.choice(len(a), n, replace=False)
            x, y = a[sel].copy(), b[sel].copy()
            if trial % 2:
                y[:2] += rng.normal(0, 5, (2, 2)).astype(np.float32)
            F0, m0 = cv2.findFundamentalMat(x, y, cv2.FM_RANSAC, 1.0, 0.99)
            F, m, ni = fe.findFundamentalMat(x, y, 1.0, 0.99)
            eq = m0 is not None and np.array_equal(m, m0.ravel())
            total += 1
            same += int(eq)
            if n == 14:            # the one size whose median is not taken among the exact-fit sample points
                assert eq, (n, trial)
                assert np.allclose(F / F[2, 2], F0 / F0[2, 2], rtol=1e-9, atol=1e-12)
    # N <= 13: the median is taken among the seven sample points, which fit their own model exactly -- it is
    # rounding residue (~1e-27) and decides the winner, in OpenCV too.  The library follows the procedure (same samples,
    # same median rule, same sigma); WHICH of the exact-fit models wins then depends on the last bits of the 7-point
    # solutions, so agreement below 14 points is statistical (oracle/replay.fmat_lmeds with cv2's own models: 40/42).
    assert same >= 7, (same, total)          # at least the six n == 14 cases and some of the others
    s7 = rng.choice(len(a), 7, replace=False)     # (the first grid points are collinear: OpenCV asserts on those)
    F0, m0 = cv2.findFundamentalMat(a[s7], b[s7], cv2.FM_RANSAC, 1.0, 0.99)
    F, m, ni = fe.findFundamentalMat(a[s7], b[s7], 1.0, 0.99)
    assert np.all(m == 1) and ni == 7
    assert np.allclose(F, F0[:3], rtol=1e-9, atol=1e-12)
    with pytest.raises(Exception):
        fe.findFundamentalMat(a[s7[:6]], b[s7[:6]], 1.0, 0.99)
    # PnP: 5 points -> EPnP on all of them, 4 points -> P3P, every point an inlier, no refinement
    from ros_stereo_slam_b200 import _lib
    for seed in range(6):
        X, xy, _, _, _ = synth.pnp_stress_case(60, 0.0, 0.2, seed=seed)
        for n in (5, 4):
            ok_, r0, t0, inl0 = cv2.solvePnPRansac(X[:n].reshape(-1, 1, 3), xy[:n].reshape(-1, 1, 2), glue.K, np.zeros((4, 1)),
                                                   None, None, False, 100, 1.0, 0.99)
            res = fe.solvePnPRansac(X[:n], xy[:n], 100, 1.0, 0.99)
            assert res["ok"] == bool(ok_)
            assert np.array_equal(res["inliers"], np.arange(n))
            tol = 1e-9 if n == 5 else 1e-5     # P3P: another formulation of the same quartic than OpenCV's (required: 1e-4 rad, 1e-3 m)
            assert np.abs(res["rvec"] - r0.ravel()).max() <= tol and np.abs(res["tvec"] - t0.ravel()).max() <= tol, (n, seed)
        res = fe.solvePnPRansac(X[:4], xy[:4], 100, 1.0, 0.99, min_solver=_lib.VO_PNP_P3P4)
        assert res["ok"] and len(res["inliers"]) == 4
        with pytest.raises(Exception):
            fe.solvePnPRansac(X[:40], xy[:40], 100, 1.0, 0.99, min_solver=_lib.VO_PNP_P3P4)   # RANSAC over P3P samples: not offered
        with pytest.raises(Exception):
            fe.solvePnPRansac(X[:3], xy[:3], 100, 1.0, 0.99)


def test_sequence_survives_a_nearly_dead_track():
    """A sequence whose reference set shrinks below 15 points: the reference carries on through OpenCV's small-N paths
    until solvePnPRansac returns fewer than 10 inliers twice (SHUTDOWN_FLAG); so must the library, frame by frame."""
    sc = synth.Scene(3)
    Ls = [sc.render(i, "L") for i in range(4)]
    Rs = [sc.render(i, "R") for i in range(4)]
    fe = make_frontend(kf_min_inliers=0)          # never insert a keyframe: the set only shrinks
    fe.seq_init(Ls[0], Rs[0])
    xy, xyz = fe.seq_reference()
    # keep 13 points of the reference set: frame 1 goes through LMedS (8..14 points), then PnP-RANSAC on what is left
    keep = np.arange(0, len(xy), max(1, len(xy) // 13))[:13]
    ref = glue.perspective_n_point_estimation(Ls[0], Ls[1], xy[keep], xyz[keep], iters=100)
    res = fe.PerspectiveNpointEstimation(Ls[0], Ls[1], xy[keep], xyz[keep])
    assert res["shutdown"] == ref["shutdown"]
    assert len(res["trk2d"]) == len(ref["trk2d"]) or len(ref["trk2d"]) < 14     # LMedS at <= 13 points is noise-decided
    if len(res["trk2d"]) == len(ref["trk2d"]) and not ref["shutdown"] and len(ref["trk2d"]) >= 6:
        assert np.array_equal(res["inliers"], ref["inliers"])
    fe.close()
