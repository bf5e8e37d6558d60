"""Host-side mirror of the reference's hot-path interface over the C ABI.

`VisualFrontEnd` exposes the member functions of the reference's `class visualSLAM` that make
up the VO front-end (reference include/visualSLAM.h:152-169) under the same names and argument
meaning, with numpy arrays standing in for cv::Mat / std::vector<Point2f|Point3f>.  Every call
goes through libvo_b200.so (hand-written CUDA for sm_100a); nothing is computed on the CPU.
The C++ equivalent for the ROS node is include/vo_b200.hpp.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VoError, VoFrameResult, check


def _f32(a, cols):
    a = np.ascontiguousarray(a, np.float32)
    if a.size == 0:
        return a.reshape(0, cols)
    return a.reshape(-1, cols)


def _u8img(img):
    if img is None:
        return None
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] != 3):
        raise ValueError("images must be uint8, H x W (gray) or H x W x 3 (BGR)")
    if img.ndim == 2 and img.strides[1] != 1:
        img = np.ascontiguousarray(img)
    if img.ndim == 3 and (img.strides[2] != 1 or img.strides[1] != 3):
        img = np.ascontiguousarray(img)
    return img


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class VisualFrontEnd:
    def __init__(self, **params):
        self.lib = _lib.load()
        self.params = _lib.default_params(**params)
        h = C.c_void_p()
        check(self.lib.vo_create(C.byref(self.params), C.byref(h)))
        self.h = h
        self.cap = self.params.max_points

    def close(self):
        if getattr(self, "h", None):
            self.lib.vo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ image marshalling
    def _ctx_img(self, img):
        """An image of the CONTEXT's geometry (the C ABI takes width/height/channels from vo_params, not from the
        call): wrong shapes or channel counts would be read with the wrong layout, so they are rejected here."""
        a = _u8img(img)
        p = self.params
        cn = 1 if a.ndim == 2 else a.shape[2]
        if a.shape[0] != p.height or a.shape[1] != p.width or cn != p.channels:
            raise ValueError("image is %s, the context was created for %d x %d x %d" % (a.shape, p.height, p.width, p.channels))
        return a

    def _ctx_pair(self, img_a, img_b):
        """Two context-geometry images behind ONE row stride (the two-image entry points take a single stride)."""
        a, b = self._ctx_img(img_a), self._ctx_img(img_b)
        if a.strides[0] != b.strides[0]:
            a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        return a, b

    # ------------------------------------------------------------------ intrinsics helpers
    @property
    def K(self):
        p = self.params
        return np.array([[p.fx, 0, p.cx], [0, p.fy, p.cy], [0, 0, 1]], np.float64)

    # ------------------------------------------------------------------ a-1 .. a-9 raw stages
    def denseKeypointExtractor(self, img, stepSize):
        """visualSLAM::denseKeypointExtractor (src/tracking.cpp:4-12) -> (N,2) float32."""
        rows, cols = img.shape[:2]
        out = np.empty((self.cap, 2), np.float32)
        n = C.c_int()
        check(self.lib.vo_grid_keypoints(self.h, rows, cols, int(stepSize), _p(out), self.cap, C.byref(n)))
        return out[:n.value].copy()

    def adaptiveNonMaximalSuppresion(self, xy, response, numToKeep):
        """adaptiveNonMaximalSuppresion (src/ANMS.cpp:18-67) -> kept indices, canonical order."""
        xy = _f32(xy, 2)
        resp = np.ascontiguousarray(response, np.float32)
        keep = np.empty(max(len(xy), 1), np.int32)
        n = C.c_int()
        check(self.lib.vo_anms(self.h, _p(xy), _p(resp), len(xy), int(numToKeep), _p(keep), len(keep), C.byref(n)))
        return keep[:n.value].copy()

    def calcOpticalFlowPyrLK(self, prevImg, nextImg, prevPts):
        """cv::calcOpticalFlowPyrLK with the reference's defaults -> (nextPts, status, err)."""
        a, b = self._ctx_pair(prevImg, nextImg)
        pts = _f32(prevPts, 2)
        n = len(pts)
        nxt = np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        check(self.lib.vo_lk_track(self.h, _p(a), _p(b), a.strides[0], _p(pts), n, _p(nxt), _p(st), _p(err)))
        return nxt, st, err

    def pyramid_level(self, img, level):
        a = self._ctx_img(img)
        w, h = C.c_int(), C.c_int()
        check(self.lib.vo_debug_pyramid_level(self.h, _p(a), a.strides[0], level, None, None, C.byref(w), C.byref(h)))
        cn = self.params.channels
        lv = np.zeros((h.value, w.value) if cn == 1 else (h.value, w.value, cn), np.uint8)
        dv = np.zeros((h.value, w.value, 2 * cn), np.int16)
        check(self.lib.vo_debug_pyramid_level(self.h, _p(a), a.strides[0], level, _p(lv), _p(dv), C.byref(w), C.byref(h)))
        return lv, dv

    def SORcloud(self, ref3d, colorMap=None, mean_k=200, stddev_mul=0.01, return_distances=False):
        """visualSLAM::SORcloud (reference src/rosFuncs.cpp:9-39): returns the filtered points (and colours);
        with return_distances also the kept indices and the per-point mean neighbour distances."""
        pts = _f32(ref3d, 3)
        n = len(pts)
        keep = np.zeros(max(n, 1), np.int32)
        nk = C.c_int()
        md = np.zeros(max(n, 1), np.float32)
        check(self.lib.vo_sor_cloud(self.h, _p(pts), n, int(mean_k), C.c_double(stddev_mul), _p(keep), n, C.byref(nk), _p(md)))
        idx = keep[:nk.value].copy()
        out = (pts[idx], None if colorMap is None else np.asarray(colorMap)[idx])
        if return_distances:
            return out + (idx, md[:n].copy())
        return out

    # ---- dense stereo (reference class StereoProcess, include/stereoCV.h:63-64)
    def sgbm_params(self, **kw):
        """vo_sgbm_params with the reference's StereoSGBM::create arguments (src/StereoCV.cpp:39-50)."""
        p = _lib.VoSgbmParams()
        self.lib.vo_sgbm_default_params(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, int(v))
        return p

    def stereoMatch(self, imL, imR, download=True, **kw):
        """StereoProcess::stereoMatch (src/StereoCV.cpp:21-62): BGR frames are converted to gray on the device,
        gray frames are taken as they are; returns the int16 16x disparity of StereoSGBM::compute.  Keyword
        arguments override the reference's SGBM parameters (names of vo_sgbm_params)."""
        a, b = _u8img(imL), _u8img(imR)
        if a.shape != b.shape:
            raise ValueError("left / right shapes differ")
        if a.strides[0] != b.strides[0]:
            a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        h, w = a.shape[:2]
        p = self.sgbm_params(**kw)
        out = np.zeros((h, w), np.int16) if download else None
        fn = self.lib.vo_stereo_match if a.ndim == 3 else self.lib.vo_sgbm_compute
        check(fn(self.h, _p(a), _p(b), a.strides[0], w, h, C.byref(p), _p(out), 2 * w))
        return out

    def reprojectDisparity(self, disp, Q, shape=None, cap=None):
        """StereoProcess::reprojectDisparity (src/StereoCV.cpp:221-250): (points (M,3) float32, flat pixel indices
        (M,)) in raster order.  disp=None reprojects the device-resident result of the last stereoMatch (pass its
        shape)."""
        Q = np.ascontiguousarray(Q, np.float64).reshape(4, 4)
        if disp is not None:
            disp = np.ascontiguousarray(disp, np.int16)
            h, w = disp.shape
        else:
            h, w = shape
        cap = h * w if cap is None else int(cap)
        xyz = np.zeros((max(cap, 1), 3), np.float32)
        idx = np.zeros(max(cap, 1), np.int32)
        n = C.c_int()
        check(self.lib.vo_reproject_disparity(self.h, _p(disp), 2 * w, w, h, _p(Q), _p(xyz), _p(idx), cap, C.byref(n)),
              ok=(_lib.VO_OK, _lib.VO_ERR_CAPACITY))
        m = min(n.value, cap)
        return xyz[:m].copy(), idx[:m].copy()

    def sgbm_timing(self):
        """Device milliseconds of the last SGBM call; the per-stage entries are filled only for calls made while
        profiling is enabled (plain launches instead of the CUDA graph), 'pipeline' always."""
        ms = (C.c_float * 9)()
        tot = C.c_float()
        check(self.lib.vo_sgbm_timing(self.h, ms, C.byref(tot)))
        names = ("upload", "prefilter", "cost_volume", "paths_vertical", "paths_other_tail", "wta",
                 "lrcheck_median", "speckle", "download")
        out = dict(zip(names, [float(v) for v in ms]))
        out["pipeline"] = float(tot.value)
        return out

    def sgbm_stage(self, stage, shape, dtype):
        out = np.zeros(shape, dtype)
        check(self.lib.vo_debug_sgbm_stage(self.h, int(stage), _p(out), C.c_uint64(out.nbytes)))
        return out

    # ---- ORB descriptor stage (reference src/optimizationStuff.cpp:49-56)
    def orbDescribe(self, img, xy, angle_deg=None):
        """rBRIEF descriptors (n x 32 uint8) of caller-made keypoints on one pyramid level, = OpenCV's ORB::compute;
        angle_deg=None computes the intensity-centroid angles first, as detectAndCompute does."""
        a = _u8img(img)
        if a.ndim != 2:
            raise ValueError("gray image expected")
        pts = _f32(xy, 2)
        ang = None if angle_deg is None else np.ascontiguousarray(angle_deg, np.float32).reshape(-1)
        n = len(pts)
        desc = np.zeros((max(n, 1), 32), np.uint8)
        check(self.lib.vo_orb_describe(self.h, _p(a), a.strides[0], a.shape[1], a.shape[0], _p(pts), _p(ang), n, _p(desc)))
        return desc[:n]

    def orbAngles(self, img, xy):
        """ORB's orientation step (ICAngles): angle in degrees of each keypoint."""
        a = _u8img(img)
        pts = _f32(xy, 2)
        n = len(pts)
        ang = np.zeros(max(n, 1), np.float32)
        check(self.lib.vo_orb_angles(self.h, _p(a), a.strides[0], a.shape[1], a.shape[0], _p(pts), n, _p(ang)))
        return ang[:n]

    def orbDetectAndCompute(self, img, nfeatures=500):
        """ORB::create(nfeatures)->detectAndCompute (reference src/optimizationStuff.cpp:49-56): dict of xy, octave,
        response, angle, desc sorted by (octave, y, x)."""
        a = _u8img(img)
        cap = 2 * int(nfeatures) + 4096
        xy = np.zeros((cap, 2), np.float32)
        octv = np.zeros(cap, np.int32)
        resp = np.zeros(cap, np.float32)
        ang = np.zeros(cap, np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int()
        check(self.lib.vo_orb_detect_and_compute(self.h, _p(a), a.strides[0], a.shape[1], a.shape[0], int(nfeatures), _p(xy),
                                                 _p(octv), _p(resp), _p(ang), _p(desc), cap, C.byref(n)))
        m = n.value
        return dict(xy=xy[:m].copy(), octave=octv[:m].copy(), response=resp[:m].copy(), angle=ang[:m].copy(),
                    desc=desc[:m].copy())

    def fast9(self, img, threshold=20, nonmax=True, cap=200000):
        """cv::FAST (TYPE_9_16): (xy (n x 2), score (n,)) in raster order."""
        a = _u8img(img)
        xy = np.zeros((cap, 2), np.float32)
        sc = np.zeros(cap, np.float32)
        n = C.c_int()
        check(self.lib.vo_fast9(self.h, _p(a), a.strides[0], a.shape[1], a.shape[0], int(threshold), int(bool(nonmax)),
                                _p(xy), _p(sc), cap, C.byref(n)), ok=(_lib.VO_OK, _lib.VO_ERR_CAPACITY))
        m = min(n.value, cap)
        return xy[:m].copy(), sc[:m].copy()

    def orbHarris(self, img, xy):
        """The Harris response ORB ranks its keypoints by (HarrisResponses, block 7, k 0.04)."""
        a = _u8img(img)
        pts = _f32(xy, 2)
        n = len(pts)
        r = np.zeros(max(n, 1), np.float32)
        check(self.lib.vo_orb_harris(self.h, _p(a), a.strides[0], a.shape[1], a.shape[0], _p(pts), n, _p(r)))
        return r[:n]

    def orbSmooth(self, img):
        a = _u8img(img)
        out = np.zeros(a.shape, np.uint8)
        check(self.lib.vo_orb_smooth(self.h, _p(a), a.strides[0], a.shape[1], a.shape[0], _p(out), out.strides[0]))
        return out

    def cvtColorBGR2GRAY(self, bgr):
        """cv::cvtColor(bgr, CV_BGR2GRAY) on the device (bit-exact with OpenCV's 15-bit fixed point)."""
        a = np.ascontiguousarray(bgr)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("expected an H x W x 3 uint8 image")
        out = np.zeros(a.shape[:2], np.uint8)
        check(self.lib.vo_bgr_to_gray(self.h, _p(a), a.strides[0], 0, _p(out), out.strides[0]))
        return out

    def pyramid_padded(self, img, level, pad=21):
        """Level `level` and its derivative with `pad` border pixels, as the LK window sees them."""
        a = _u8img(img)
        w, h = C.c_int(), C.c_int()
        check(self.lib.vo_debug_pyramid_level(self.h, _p(a), a.strides[0], level, None, None, C.byref(w), C.byref(h)))
        cn = self.params.channels
        lv = np.zeros((h.value + 2 * pad, w.value + 2 * pad) if cn == 1 else (h.value + 2 * pad, w.value + 2 * pad, cn), np.uint8)
        dv = np.zeros((h.value + 2 * pad, w.value + 2 * pad, 2 * cn), np.int16)
        check(self.lib.vo_debug_pyramid_padded(self.h, _p(a), a.strides[0], level, pad, _p(lv), _p(dv)))
        return lv, dv

    def findFundamentalMat(self, pts1, pts2, thr, conf=0.99, samples=None):
        """cv::findFundamentalMat(FM_RANSAC) -> (F 3x3 or None, mask (N,) u8, n_inliers)."""
        a, b = _f32(pts1, 2), _f32(pts2, 2)
        n = len(a)
        mask = np.zeros(n, np.uint8)
        F = np.zeros(9, np.float64)
        ni = C.c_int()
        s = None if samples is None else np.ascontiguousarray(samples, np.int32).reshape(-1, 7)
        r = self.lib.vo_fmat_ransac(self.h, _p(a), _p(b), n, C.c_double(thr), C.c_double(conf),
                                    _p(s), 0 if s is None else len(s), _p(mask), _p(F), C.byref(ni))
        check(r, (_lib.VO_OK, _lib.VO_ERR_NO_MODEL))
        return (F.reshape(3, 3) if r == _lib.VO_OK else None), mask, ni.value

    def last_fmat(self):
        cap = self.params.max_hypotheses
        models = np.zeros((cap, 3, 9), np.float64)
        counts = np.zeros((cap, 3), np.int32)
        nh, bs, bm, nit = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self.lib.vo_debug_last_fmat(self.h, _p(models), _p(counts), cap, C.byref(nh), C.byref(bs), C.byref(bm),
                                          C.byref(nit)))
        return dict(models=models[:nh.value], counts=counts[:nh.value], best=(bs.value, bm.value), n_iters=nit.value)

    def triangulatePoints(self, P1, P2, pts1, pts2):
        """cv::triangulatePoints + float dehomogenisation (src/triangulation.cpp:152-160)."""
        a, b = _f32(pts1, 2), _f32(pts2, 2)
        P1 = np.ascontiguousarray(P1, np.float64)
        P2 = np.ascontiguousarray(P2, np.float64)
        out = np.zeros((len(a), 3), np.float32)
        check(self.lib.vo_triangulate(self.h, _p(P1), _p(P2), _p(a), _p(b), len(a), _p(out)))
        return out

    def solvePnPRansac(self, pts3d, pts2d, iters=100, thr=1.0, conf=0.99, samples=None, min_solver=None):
        """cv::solvePnPRansac(..., false, iters, thr, conf, inliers) -> dict(ok, rvec, tvec, inliers)."""
        X, x = _f32(pts3d, 3), _f32(pts2d, 2)
        n = len(X)
        rvec = np.zeros(3)
        tvec = np.zeros(3)
        inl = np.zeros(max(n, 1), np.int32)
        ni = C.c_int()
        s = None if samples is None else np.ascontiguousarray(samples, np.int32).reshape(-1, 5)
        r = self.lib.vo_pnp_ransac(self.h, _p(X), _p(x), n, int(iters), C.c_double(thr), C.c_double(conf),
                                   _lib.VO_PNP_EPNP5 if min_solver is None else int(min_solver), _p(s), 0 if s is None else len(s),
                                   _p(rvec), _p(tvec), _p(inl), len(inl), C.byref(ni))
        check(r, (_lib.VO_OK, _lib.VO_ERR_NO_MODEL))
        return dict(ok=r == _lib.VO_OK, rvec=rvec, tvec=tvec, inliers=inl[:ni.value].copy())

    def last_pnp(self):
        cap = self.params.max_hypotheses
        models = np.zeros((cap, 6), np.float64)
        counts = np.zeros(cap, np.int32)
        nh, best, nit = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.vo_debug_last_pnp(self.h, _p(models), _p(counts), cap, C.byref(nh), C.byref(best), C.byref(nit)))
        return dict(models=models[:nh.value], counts=counts[:nh.value], best=best.value, n_iters=nit.value)

    def update3dtransformation(self, pts3d, pose3x4):
        """visualSLAM::update3dtransformation (src/keyFrameManagement.cpp:33-46)."""
        X = _f32(pts3d, 3)
        M = np.ascontiguousarray(pose3x4, np.float64).reshape(3, 4)
        out = np.zeros_like(X)
        check(self.lib.vo_transform_points(self.h, _p(M), _p(X), len(X), _p(out)))
        return out

    def pose_from_pnp(self, rvec, tvec):
        """src/VisualSLAM.cpp:70-74,93-97 -> 3x4 [R|t] camera->world."""
        r = np.ascontiguousarray(rvec, np.float64).ravel()
        t = np.ascontiguousarray(tvec, np.float64).ravel()
        M = np.zeros(12)
        check(self.lib.vo_pose_from_pnp(_p(r), _p(t), _p(M)))
        return M.reshape(3, 4)

    # ------------------------------------------------------------------ the reference's method boundaries
    def denseLKtracking(self, refImg, curImg, refPts):
        """visualSLAM::denseLKtracking (src/tracking.cpp:14-28) -> (refPts_kept, trackPts_kept)."""
        a, b = self._ctx_pair(refImg, curImg)
        pts = _f32(refPts, 2)
        n = len(pts)
        r = np.zeros((n, 2), np.float32)
        t = np.zeros((n, 2), np.float32)
        m = C.c_int()
        check(self.lib.vo_dense_lk_tracking(self.h, _p(a), _p(b), a.strides[0], _p(pts), n, _p(r), _p(t), C.byref(m)))
        return r[:m.value].copy(), t[:m.value].copy()

    def FmatThresholding(self, refPts, trkPts):
        """visualSLAM::FmatThresholding (src/tracking.cpp:30-43)."""
        a, b = _f32(refPts, 2), _f32(trkPts, 2)
        n = len(a)
        r = np.zeros((n, 2), np.float32)
        t = np.zeros((n, 2), np.float32)
        m = C.c_int()
        check(self.lib.vo_fmat_thresholding(self.h, _p(a), _p(b), n, _p(r), _p(t), C.byref(m)))
        return r[:m.value].copy(), t[:m.value].copy()

    def stereoTriangulate(self, im1, im2):
        """visualSLAM::stereoTriangulate (src/triangulation.cpp:73-166) -> (ref3dPts, ref2dPts)."""
        a, b = self._ctx_pair(im1, im2)
        if a is None or b is None:
            print("NULL IMG")
            return None
        xyz = np.zeros((self.cap, 3), np.float32)
        xy = np.zeros((self.cap, 2), np.float32)
        n = C.c_int()
        check(self.lib.vo_stereo_triangulate(self.h, _p(a), _p(b), a.strides[0], _p(xyz), _p(xy), self.cap, C.byref(n)))
        return xyz[:n.value].copy(), xy[:n.value].copy()

    def insertKeyFrames(self, imL, imR, pose4dTransform):
        """visualSLAM::insertKeyFrames (src/keyFrameManagement.cpp:9-31)
        -> (ref3dCoords world, ftrPts, untransformed)."""
        a, b = self._ctx_pair(imL, imR)
        M = np.ascontiguousarray(pose4dTransform, np.float64).reshape(3, 4)
        xyz = np.zeros((self.cap, 3), np.float32)
        cam = np.zeros((self.cap, 3), np.float32)
        xy = np.zeros((self.cap, 2), np.float32)
        n = C.c_int()
        check(self.lib.vo_insert_keyframe(self.h, _p(a), _p(b), a.strides[0], _p(M), _p(xyz), _p(xy), _p(cam), self.cap,
                                          C.byref(n)))
        return xyz[:n.value].copy(), xy[:n.value].copy(), cam[:n.value].copy()

    def PyrLKtrackFrame2Frame(self, refimg, curImg, refPts, ref3dpts):
        """visualSLAM::PyrLKtrackFrame2Frame (src/tracking.cpp:46-91)
        -> (refRetpts = tracked 2-D, ref3dretPts, inlierReferencePyrLKPts)."""
        a, b = self._ctx_pair(refimg, curImg)
        p2, p3 = _f32(refPts, 2), _f32(ref3dpts, 3)
        n = len(p2)
        t2 = np.zeros((n, 2), np.float32)
        t3 = np.zeros((n, 3), np.float32)
        r2 = np.zeros((n, 2), np.float32)
        m = C.c_int()
        check(self.lib.vo_track_frame(self.h, _p(a), _p(b), a.strides[0], _p(p2), _p(p3), n, _p(t2), _p(t3), _p(r2),
                                      C.byref(m)))
        return t2[:m.value].copy(), t3[:m.value].copy(), r2[:m.value].copy()

    def PerspectiveNpointEstimation(self, prevImg, curImg, ref2dPoints, ref3dPoints):
        """visualSLAM::PerspectiveNpointEstimation (src/keyFrameManagement.cpp:73-94)
        -> dict(trk2d, trk3d, ref2d_inl, rvec, tvec, inliers, attempt, shutdown)."""
        a, b = self._ctx_pair(prevImg, curImg)
        p2, p3 = _f32(ref2dPoints, 2), _f32(ref3dPoints, 3)
        n = len(p2)
        t2 = np.zeros((n, 2), np.float32)
        t3 = np.zeros((n, 3), np.float32)
        r2 = np.zeros((n, 2), np.float32)
        rvec, tvec = np.zeros(3), np.zeros(3)
        inl = np.zeros(max(n, 1), np.int32)
        m, ni, att = C.c_int(), C.c_int(), C.c_int()
        r = self.lib.vo_pnp_frame(self.h, _p(a), _p(b), a.strides[0], _p(p2), _p(p3), n, _p(t2), _p(t3), _p(r2),
                                  C.byref(m), _p(rvec), _p(tvec), _p(inl), len(inl), C.byref(ni), C.byref(att))
        check(r, (_lib.VO_OK, _lib.VO_ERR_LOW_INLIERS))
        return dict(trk2d=t2[:m.value].copy(), trk3d=t3[:m.value].copy(), ref2d_inl=r2[:m.value].copy(), rvec=rvec,
                    tvec=tvec, inliers=inl[:ni.value].copy(), attempt=att.value,
                    shutdown=(r == _lib.VO_ERR_LOW_INLIERS))

    # ------------------------------------------------------------------ device-resident sequence driver
    def seq_init(self, left, right, is_device=False, stride=None):
        n = C.c_int()
        if is_device:
            check(self.lib.vo_seq_init(self.h, C.c_void_p(left), C.c_void_p(right), stride or self.params.width, 1,
                                       C.byref(n)))
        else:
            a, b = self._ctx_pair(left, right)
            check(self.lib.vo_seq_init(self.h, _p(a), _p(b), a.strides[0], 0, C.byref(n)))
        return n.value

    def seq_prefetch(self, left, right=None):
        """Announce the next frame's host images (vo_seq_prefetch); pass the SAME arrays to seq_track later."""
        a, b = self._ctx_pair(left, right)
        self._prefetched = (a, b)            # keep them alive until they are consumed
        check(self.lib.vo_seq_prefetch(self.h, _p(a), _p(b), a.strides[0]))

    def seq_announce(self, left, right, stride=None):
        """Announce the next frame's DEVICE-resident images (vo_seq_announce); pass the same pointers to seq_track."""
        check(self.lib.vo_seq_announce(self.h, C.c_void_p(left), C.c_void_p(right) if right else None,
                                       stride or self.params.width * self.params.channels, 1))

    def seq_track(self, left, right=None, is_device=False, stride=None, force_keyframe=False):
        res = VoFrameResult()
        if is_device:
            r = self.lib.vo_seq_track(self.h, C.c_void_p(left), C.c_void_p(right) if right else None,
                                      stride or self.params.width, 1, int(force_keyframe), C.byref(res))
        else:
            a, b = self._ctx_pair(left, right)
            r = self.lib.vo_seq_track(self.h, _p(a), _p(b), a.strides[0], 0, int(force_keyframe), C.byref(res))
        check(r, (_lib.VO_OK, _lib.VO_ERR_LOW_INLIERS))
        return res, r

    def seq_reference(self):
        xy = np.zeros((self.cap, 2), np.float32)
        xyz = np.zeros((self.cap, 3), np.float32)
        n = C.c_int()
        check(self.lib.vo_seq_get_reference(self.h, _p(xy), _p(xyz), self.cap, C.byref(n)))
        return xy[:n.value].copy(), xyz[:n.value].copy()

    # ------------------------------------------------------------------ measurement helpers
    def sync(self):
        check(self.lib.vo_sync(self.h))

    def profile_enable(self, kinds="all"):
        """kinds: "all", None/False (off) or an iterable of names from _lib.KERNELS."""
        if kinds in (None, False, 0):
            mask = 0
        elif kinds == "all" or kinds is True:
            mask = -1
        else:
            mask = 0
            for k in kinds:
                mask |= 1 << _lib.KERNELS.index(k)
        check(self.lib.vo_profile_enable(self.h, C.c_int(mask)))

    def profile_read(self, reset=False):
        out = {}
        for i, name in enumerate(_lib.KERNELS):
            l, ms = C.c_int64(), C.c_double()
            check(self.lib.vo_profile_read(self.h, i, C.byref(l), C.byref(ms), int(reset)))
            out[name] = (l.value, ms.value)
        return out

    def launch_count(self):
        return self.lib.vo_launch_count(self.h)

    def lk_work(self):
        a, b = C.c_int64(), C.c_int64()
        check(self.lib.vo_lk_work(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def lk_slow_paths(self):
        """(window extractions, iterations) that took the sequential float-chain path since vo_create."""
        a, b = C.c_int64(), C.c_int64()
        check(self.lib.vo_lk_slow_paths(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_fp32_peak(self):
        t = C.c_double()
        check(self.lib.vo_measure_fp32_peak(self.h, C.byref(t)))
        return t.value

    def measure_int32_peak(self):
        t = C.c_double()
        check(self.lib.vo_measure_int32_peak(self.h, C.byref(t)))
        return t.value

    def synth_render(self, seed, frame, eye):
        """Harness: render one synthetic frame on the GPU and return it as a numpy array."""
        w, h = self.params.width, self.params.height
        d = C.c_void_p()
        check(self.lib.vo_alloc_dev(self.h, C.byref(d), w * h))
        try:
            check(self.lib.vo_synth_render_dev(self.h, seed, frame, eye, d))
            out = np.zeros((h, w), np.uint8)
            check(self.lib.vo_memcpy_d2h(self.h, _p(out), d, w * h))
        finally:
            self.lib.vo_free_dev(self.h, d)
        return out
