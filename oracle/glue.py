"""Call-through restatement of the reference's hot-path glue.  TEST INFRASTRUCTURE ONLY.

Every function follows one reference method and calls the same OpenCV entry
point with the reference's arguments (file:line cited per function, paths are
into the reference tree).  Points are numpy float32 arrays (N,2)/(N,3) instead
of std::vector<Point2f/Point3f>.  This is the parity target ("reference
arithmetic") and the CPU baseline timed by bench.py.
"""
import numpy as np
import cv2

# include/visualSLAM.h:68,82-87
FX = 7.188560000000e+02
FY = 7.188560000000e+02
CX = 6.071928000000e+02
CY = 1.852157000000e+02
BASELINE = 0.54
K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], np.float64)


def dense_keypoint_extractor(rows, cols, step):
    """visualSLAM::denseKeypointExtractor, src/tracking.cpp:4-12.
    Raster grid, y-major, x fastest; returns (N,2) float32 of (x,y)."""
    ys = np.arange(step, rows - step, step, dtype=np.float32)
    xs = np.arange(step, cols - step, step, dtype=np.float32)
    if len(ys) == 0 or len(xs) == 0:
        return np.zeros((0, 2), np.float32)
    gx, gy = np.meshgrid(xs, ys)
    return np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)


def anms(xy, response, num_to_keep):
    """adaptiveNonMaximalSuppresion, src/ANMS.cpp:18-67.

    Returns the indices (into the input) of the kept keypoints in the canonical
    order (response desc, original index asc).  std::sort at ANMS.cpp:26 is
    unstable, so only the kept SET is defined by the reference; the canonical
    order is ours.  ``size < numToKeep`` returns
    everything (ANMS.cpp:21); ``size == numToKeep`` reads out of bounds in the
    reference (ANMS.cpp:59) and is rejected here.
    """
    n = len(xy)
    if n < num_to_keep:
        return np.arange(n, dtype=np.int32)
    if n == num_to_keep:
        raise ValueError("ANMS.cpp:59 reads radiiSorted[numToKeep] out of bounds when size == numToKeep")
    order = np.lexsort((np.arange(n), -response.astype(np.float64)))
    pts = xy[order].astype(np.float32)
    resp = response[order].astype(np.float32)
    radii = np.full(n, np.finfo(np.float64).max)
    robust = np.float32(1.11)
    for i in range(n):
        r = np.float32(resp[i] * robust)
        # loop stops at the first j whose response is not > r (ANMS.cpp:45);
        # in sorted order that is a prefix.
        stop = i
        fail = np.nonzero(~(resp[:i] > r))[0]
        if len(fail):
            stop = fail[0]
        if stop > 0:
            d = (pts[i] - pts[:stop]).astype(np.float32)  # Point2f subtraction is float
            dist = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2)
            radii[i] = dist.min()
    decision = np.sort(radii)[::-1][num_to_keep]
    keep = radii >= decision
    return order[keep].astype(np.int32)


def dense_lk_tracking(ref_img, cur_img, ref_pts):
    """visualSLAM::denseLKtracking, src/tracking.cpp:14-28.
    Returns (ref_pts_kept, trk_pts_kept) = status==1 subsets, order preserved."""
    if len(ref_pts) == 0:
        return ref_pts.copy(), ref_pts.copy()
    trk, status, _err = cv2.calcOpticalFlowPyrLK(ref_img, cur_img, ref_pts.reshape(-1, 1, 2), None)
    keep = status.ravel() == 1
    return ref_pts[keep].copy(), trk.reshape(-1, 2)[keep].copy()


def fmat_thresholding(ref_pts, trk_pts, thr=3.0, conf=0.99):
    """visualSLAM::FmatThresholding, src/tracking.cpp:30-43 (CV_RANSAC, 3.0, 0.99)."""
    F, mask = cv2.findFundamentalMat(ref_pts, trk_pts, cv2.FM_RANSAC, thr, conf)
    if mask is None:
        # reference would walk an empty mask: nothing survives
        return ref_pts[:0].copy(), trk_pts[:0].copy(), np.zeros(0, np.uint8)
    keep = mask.ravel() == 1
    return ref_pts[keep].copy(), trk_pts[keep].copy(), mask.ravel().copy()


def pyr_lk_track_frame2frame(ref_img, cur_img, ref_pts, ref_3d, thr=1.0, conf=0.99):
    """visualSLAM::PyrLKtrackFrame2Frame, src/tracking.cpp:46-91.

    Returns (trk2d, trk3d, ref2d_inl): the F-inlier tracked points, their 3-D
    points and the surviving reference 2-D points (inlierReferencePyrLKPts,
    tracking.cpp:90).  The reference's second loop runs to refPts.size() and
    indexes inIdx out of range when any status==0 (tracking.cpp:78-84); the
    replacement iterates inIdx.size().
    """
    trk, status, _err = cv2.calcOpticalFlowPyrLK(ref_img, cur_img, ref_pts.reshape(-1, 1, 2), None)
    keep = status.ravel() == 1
    r2 = ref_pts[keep]
    r3 = ref_3d[keep]
    t2 = trk.reshape(-1, 2)[keep]
    F, mask = cv2.findFundamentalMat(r2, t2, 8, thr, conf)
    if mask is None:
        return t2[:0].copy(), r3[:0].copy(), r2[:0].copy()
    m = mask.ravel() == 1
    return t2[m].copy(), r3[m].copy(), r2[m].copy()


def projection_matrices(k=K, baseline=BASELINE):
    """src/triangulation.cpp:142-149."""
    P1 = np.zeros((3, 4))
    P2 = np.zeros((3, 4))
    P1[0, 0] = P1[1, 1] = P1[2, 2] = 1
    P2[0, 0] = P2[1, 1] = P2[2, 2] = 1
    P2[0, 3] = -baseline
    return k @ P1, k @ P2


def triangulate(P1, P2, pt1, pt2):
    """cv::triangulatePoints + float dehomogenisation, src/triangulation.cpp:152-160."""
    if len(pt1) == 0:
        return np.zeros((0, 3), np.float32)
    est = cv2.triangulatePoints(P1, P2, pt1.T.astype(np.float32), pt2.T.astype(np.float32))
    est = est.astype(np.float32)
    return np.stack([est[0] / est[3], est[1] / est[3], est[2] / est[3]], 1).astype(np.float32)


def stereo_triangulate(im_l, im_r, step=30, thr=3.0, conf=0.99, k=K, baseline=BASELINE):
    """visualSLAM::stereoTriangulate (DENSE_FLAG branch), src/triangulation.cpp:73-166.
    Returns (ref3d (N,3) f32 camera frame, ref2d (N,2) f32 left-image points)."""
    if im_l is None or im_r is None:
        return None
    ref = dense_keypoint_extractor(im_l.shape[0], im_l.shape[1], step)
    ref, trk = dense_lk_tracking(im_l, im_r, ref)
    ref, trk, _ = fmat_thresholding(ref, trk, thr, conf)
    P1, P2 = projection_matrices(k, baseline)
    xyz = triangulate(P1, P2, ref, trk)
    return xyz, ref


def update_3d_transformation(pts3d, pose3x4):
    """visualSLAM::update3dtransformation / insertKeyFrames inner loop,
    src/keyFrameManagement.cpp:20-30,33-46: double mul-add of a float point,
    stored as float."""
    p = pts3d.astype(np.float64)
    M = np.asarray(pose3x4, np.float64)
    out = np.empty_like(p)
    for r in range(3):
        # exact operation order of the reference expression (left to right)
        out[:, r] = ((M[r, 0] * p[:, 0] + M[r, 1] * p[:, 1]) + M[r, 2] * p[:, 2]) + M[r, 3]
    return out.astype(np.float32)


def insert_key_frames(im_l, im_r, pose3x4, step=30, thr=3.0, conf=0.99):
    """visualSLAM::insertKeyFrames, src/keyFrameManagement.cpp:9-31.
    Returns (ref3d_world, ref2d, untransformed)."""
    xyz, ref2d = stereo_triangulate(im_l, im_r, step, thr, conf)
    return update_3d_transformation(xyz, pose3x4), ref2d, xyz


def perspective_n_point_estimation(ref_img, cur_img, ref2d, ref3d,
                                   iters=100, thr=1.0, conf=0.99,
                                   retry_iters=100, retry_thr=8.0, retry_conf=0.98,
                                   min_inliers=10, f_thr=1.0, f_conf=0.99):
    """visualSLAM::PerspectiveNpointEstimation, src/keyFrameManagement.cpp:73-94.

    Returns dict(trk2d, trk3d, ref2d_inl, rvec, tvec, inliers, attempt, shutdown).
    attempt = 1 or 2; shutdown mirrors SHUTDOWN_FLAG (keyFrameManagement.cpp:89-92).
    """
    trk2d, trk3d, ref_inl = pyr_lk_track_frame2frame(ref_img, cur_img, ref2d, ref3d, f_thr, f_conf)
    dist = np.zeros((4, 1))
    out = dict(trk2d=trk2d, trk3d=trk3d, ref2d_inl=ref_inl, rvec=None, tvec=None,
               inliers=np.zeros(0, np.int32), attempt=1, shutdown=False)

    def run(it, th, cf):
        if len(trk3d) < 4:
            return None, None, np.zeros(0, np.int32)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(trk3d.reshape(-1, 1, 3), trk2d.reshape(-1, 1, 2), K, dist,
                                                 None, None, False, it, th, cf)
        inl = np.zeros(0, np.int32) if inl is None else inl.ravel().astype(np.int32)
        return rvec, tvec, inl

    rvec, tvec, inl = run(iters, thr, conf)
    if len(inl) < min_inliers:
        out["attempt"] = 2
        rvec2, tvec2, inl = run(retry_iters, retry_thr, retry_conf)
        if rvec2 is not None:
            rvec, tvec = rvec2, tvec2
        if len(inl) < min_inliers:
            out["shutdown"] = True
    out.update(rvec=None if rvec is None else rvec.ravel().copy(),
               tvec=None if tvec is None else tvec.ravel().copy(), inliers=inl)
    return out


def camera_pose_from_pnp(rvec, tvec):
    """src/VisualSLAM.cpp:70-74,93-97: R = Rodrigues(rvec)^T, t = -R tvec,
    pose3x4 = [R|t] (camera -> world)."""
    R, _ = cv2.Rodrigues(np.asarray(rvec, np.float64).reshape(3, 1))
    R = R.T
    t = -R @ np.asarray(tvec, np.float64).reshape(3, 1)
    return np.hstack([R, t])


def run_sequence(frames_l, frames_r, step=30, pnp_iters=100, kf_min_inliers=200, timing=None):
    """The per-frame loop of visualSLAM::initSequence restricted to the hot
    path, src/VisualSLAM.cpp:11-214 (lines 31, 64, 70-74, 93-97, 120-146, 151).

    kf_min_inliers: the keyframe rule ``inliers.size() < 200`` (VisualSLAM.cpp:120);
    pass a huge value to insert a keyframe on every frame.
    Returns a list of per-frame dicts (rvec, tvec, n_inliers, n_tracked, n_lk_in, keyframe).
    """
    import time
    xyz, ref2d = stereo_triangulate(frames_l[0], frames_r[0], step)
    ref3d = xyz
    ref_img = frames_l[0]
    out = []
    for i in range(1, len(frames_l)):
        t0 = time.perf_counter()
        cur = frames_l[i]
        res = perspective_n_point_estimation(ref_img, cur, ref2d, ref3d, iters=pnp_iters)
        rec = dict(rvec=res["rvec"], tvec=res["tvec"], n_inliers=len(res["inliers"]),
                   n_tracked=len(res["trk2d"]), n_lk_in=len(ref2d), keyframe=False,
                   shutdown=res["shutdown"])
        if res["shutdown"]:
            out.append(rec)
            break
        pose = camera_pose_from_pnp(res["rvec"], res["tvec"])
        if len(res["inliers"]) < kf_min_inliers:
            ref3d, ref2d, _untr = insert_key_frames(frames_l[i], frames_r[i], pose, step)
            rec["keyframe"] = True
            rec["n_kf_points"] = len(ref2d)
        else:
            ref3d, ref2d = res["trk3d"], res["trk2d"]
        ref_img = cur
        rec["ms"] = (time.perf_counter() - t0) * 1e3
        out.append(rec)
    return out


def bgr_to_gray(bgr):
    """cv::cvtColor(im, gray, CV_BGR2GRAY) as the reference calls it (src/StereoCV.cpp:35-36)."""
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)


def bgr_to_gray_restated(bgr):
    """The arithmetic behind it for 8-bit images: 15-bit fixed point, round to nearest
    (pinned against cv2 in tests/test_oracle_misc.py)."""
    b, g, r = (bgr[..., i].astype(np.int32) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)
