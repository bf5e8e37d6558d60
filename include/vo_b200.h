/*
 * vo_b200.h -- C ABI of libvo_b200.so, the B200-native visual-odometry front-end.
 *
 * Drop-in boundary for the per-frame hot path of Gautham-JS/ROS_Stereo_SLAM.  The
 * reference has no FFI/plugin interface: the path sits behind C++ member functions
 * of `class visualSLAM` (reference include/visualSLAM.h:152-169), called only from
 * visualSLAM::initSequence (reference src/VisualSLAM.cpp:31,64,112,123).  Each entry
 * point below cites the reference member function (file:line in the reference tree)
 * it replaces; INTEGRATION.md shows the ~30-line C++ shim that converts cv::Mat /
 * std::vector to these raw pointers inside the ROS node.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  Every function returns
 *     VO_OK (0) or a negative VO_ERR_*; nothing throws or aborts.
 *   - all pointers are HOST pointers owned by the caller unless the name ends in
 *     `_dev` / the doc says "device"; outputs have caller-provided capacity and the
 *     number written is returned through an int*.
 *   - 2-D points are interleaved float32 (x,y) = cv::Point2f; 3-D points are
 *     interleaved float32 (x,y,z) = cv::Point3f; images are 8-bit, 1 channel,
 *     row-major with `stride` bytes per row (cv::Mat::step).
 *   - results are complete (stream-synchronised) on return.
 *   - one vo_ctx per host thread / per GPU; calls on one ctx are serialised by the
 *     caller (the reference is single-threaded on this path).
 *   - there is NO CPU fallback: without a CUDA device vo_create fails with
 *     VO_ERR_NO_DEVICE.
 */
#ifndef VO_B200_H
#define VO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define VO_B200_ABI_VERSION 1

enum {
  VO_OK = 0,
  VO_ERR_INVALID_ARG = -1,
  VO_ERR_NO_DEVICE = -2,      /* no CUDA device / driver: the library never computes on the CPU */
  VO_ERR_CUDA = -3,           /* a CUDA call failed; see vo_last_error() */
  VO_ERR_CAPACITY = -4,       /* caller buffer or ctx max_points / max_hypotheses too small */
  VO_ERR_TOO_FEW_POINTS = -5, /* fewer correspondences than the minimal solver needs */
  VO_ERR_NO_MODEL = -6,       /* RANSAC never found a model with enough inliers */
  VO_ERR_LOW_INLIERS = -7,    /* both PnP attempts gave < pnp_min_inliers: the reference's
                                 SHUTDOWN_FLAG (src/keyFrameManagement.cpp:89-92) */
  VO_ERR_NOT_IMPLEMENTED = -8,
  VO_ERR_SELF_CHECK = -9      /* device numerics differ from the IEEE host evaluation of the same code */
};

/* minimal solvers for vo_pnp_ransac.  EPNP5 = what cv::solvePnPRansac uses with the reference's (default) flags:
 * RANSAC over 5-point EPnP samples for n >= 6; for n == 5 / n == 4 OpenCV solves once on all points (EPnP / P3P),
 * keeps every point as an inlier and does not refine -- so does this library.  P3P4 names that four-point solver
 * explicitly and is accepted for n == 4 only (RANSAC over P3P samples is cv::solvePnPRansac(flags = SOLVEPNP_P3P),
 * which the reference never calls: VO_ERR_INVALID_ARG). */
enum { VO_PNP_EPNP5 = 0, VO_PNP_P3P4 = 1 };

typedef struct vo_ctx vo_ctx;

/* Parameters.  vo_default_params() fills the reference's compile-time constants. */
typedef struct vo_params {
  double fx, fy, cx, cy;      /* 718.856, 718.856, 607.1928, 185.2157  (include/visualSLAM.h:82-87) */
  double baseline;            /* 0.54                                   (include/visualSLAM.h:68)    */
  int width, height;          /* 1241, 376 */
  int channels;               /* 1 = gray; 3 = interleaved BGR as cv::imread returns it (what the reference
                               * feeds calcOpticalFlowPyrLK, src/keyFrameManagement.cpp:52,64): every image
                               * argument is then H x W x 3 with stride >= 3*width bytes */
  int lk_win;                 /* 21   cv::calcOpticalFlowPyrLK defaults (src/tracking.cpp:18,52) */
  int lk_max_level;           /* 3  */
  int lk_max_iters;           /* 30 */
  double lk_eps;              /* 0.01 */
  double lk_min_eig;          /* 1e-4 */
  int grid_step;              /* 30   (src/triangulation.cpp:89) */
  double f_thr_stereo;        /* 3.0  (src/tracking.cpp:34) */
  double f_thr_temporal;      /* 1.0  (src/tracking.cpp:75) */
  double f_conf;              /* 0.99 */
  int f_max_iters;            /* 1000 (cv::findFundamentalMat default maxIters) */
  int pnp_iters;              /* 100  (src/keyFrameManagement.cpp:84) */
  double pnp_thr;             /* 1.0  */
  double pnp_conf;            /* 0.99 */
  int pnp_retry_iters;        /* 100  (src/keyFrameManagement.cpp:88) */
  double pnp_retry_thr;       /* 8.0  */
  double pnp_retry_conf;      /* 0.98 */
  int pnp_min_inliers;        /* 10   (src/keyFrameManagement.cpp:85,89) */
  int kf_min_inliers;         /* 200  keyframe rule (src/VisualSLAM.cpp:120), used by vo_seq_* only */
  int ransac_exhaustive;      /* PnP-RANSAC. 0: solve/score only the hypotheses OpenCV's adaptive stop would
                                 reach.  1: solve and score all `iters` hypotheses (the BASELINE workload,
                                 "1024 hypotheses per frame"); results are identical either way because the
                                 acceptance replay keeps the early-exit semantics. */
  int f_exhaustive;           /* same switch for the F-matrix RANSAC (f_max_iters samples); default 0 */
  int max_points;             /* capacity: keypoints per call (default 131072) */
  int max_hypotheses;         /* capacity: RANSAC samples per call (default 4096) */
  int device;                 /* CUDA device ordinal */
} vo_params;

void vo_default_params(vo_params* p);
int vo_abi_version(void);
const char* vo_last_error(void);         /* thread-local message of the last failure */
const char* vo_strerror(int code);

int vo_create(const vo_params* p, vo_ctx** out);
int vo_destroy(vo_ctx* ctx);
/* Device-vs-host bit comparison of the FP64 solvers on canned inputs (run by vo_create). */
int vo_self_check(vo_ctx* ctx);

/* ---- a-1  visualSLAM::denseKeypointExtractor(img, step)      src/tracking.cpp:4-12 ----
 * Raster grid (y-major, x fastest) over a rows x cols image; xy receives up to cap points. */
int vo_grid_keypoints(vo_ctx* ctx, int rows, int cols, int step, float* xy, int cap, int* n);

/* ---- a-2  adaptiveNonMaximalSuppresion(keypoints, numToKeep) src/ANMS.cpp:18-67 ----
 * keep_idx receives the indices (into the input) of the kept keypoints in canonical
 * order (response desc, index asc).  n < num_keep keeps everything (ANMS.cpp:21);
 * n == num_keep is rejected (the reference reads out of bounds, ANMS.cpp:59). */
int vo_anms(vo_ctx* ctx, const float* xy, const float* response, int n, int num_keep,
            int32_t* keep_idx, int cap, int* n_keep);

/* ---- a-3  cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err) with all
 * defaults, as called at src/tracking.cpp:18 and :52.  Raw outputs (no compaction);
 * err may be NULL: the sum is then skipped, but status still behaves as if err had been requested (the
 * reference always passes it, and OpenCV's err pass can clear status at level 0). */
int vo_lk_track(vo_ctx* ctx, const uint8_t* prev, const uint8_t* next, int stride,
                const float* prev_xy, int n, float* next_xy, uint8_t* status, float* err);

/* buildOpticalFlowPyramid intermediates, for parity tests: level `level` of the pyramid
 * of `img` (tight, w x h bytes) and its Scharr derivative (w x h x 2 int16); either
 * output may be NULL.  w/h receive the level size. */
int vo_debug_pyramid_level(vo_ctx* ctx, const uint8_t* img, int stride, int level,
                           uint8_t* out_level, int16_t* out_deriv, int* w, int* h);
/* Same, including `pad` (<= 21) border pixels on every side as the LK window sees them:
 * out_level is (h+2*pad) x (w+2*pad) u8 (BORDER_REFLECT_101), out_deriv (h+2*pad) x (w+2*pad)
 * x 2 int16 (zero border) -- the padded layout cv::buildOpticalFlowPyramid produces. */
int vo_debug_pyramid_padded(vo_ctx* ctx, const uint8_t* img, int stride, int level, int pad,
                            uint8_t* out_level, int16_t* out_deriv);

/* ---- a-5  cv::findFundamentalMat(pts1, pts2, FM_RANSAC, thr, conf, mask)
 * src/tracking.cpp:34 (thr 3.0) and :75 (thr 1.0).
 * samples7: NULL -> OpenCV's RNG stream (seed 0xFFFFFFFFFFFFFFFF) incl. the collinearity
 * subset check; else an n_samples x 7 replay list of accepted subsets.  mask receives n
 * bytes (0/1); F (row-major 3x3) and n_inliers may be NULL.  Below 15 points OpenCV does not run
 * RANSAC and neither does this call: 8 <= n <= 14 -> OpenCV's LMedS estimator (a fixed number of 7-point
 * samples, smallest median error, mask = err <= sigma^2; thr is not used), n == 7 -> the raw 7-point result
 * (its first solution in F) with a mask of ones, n < 7 -> VO_ERR_TOO_FEW_POINTS (OpenCV returns an empty
 * matrix and no mask). */
int vo_fmat_ransac(vo_ctx* ctx, const float* xy1, const float* xy2, int n, double thr, double conf,
                   const int32_t* samples7, int n_samples, uint8_t* mask, double F[9], int* n_inliers);

/* ---- a-6  cv::triangulatePoints(P1, P2, pt1, pt2) + float dehomogenisation
 * src/triangulation.cpp:152-160.  P1, P2 row-major 3x4. */
int vo_triangulate(vo_ctx* ctx, const double P1[12], const double P2[12],
                   const float* xy1, const float* xy2, int n, float* xyz);

/* ---- a-7  cv::solvePnPRansac(p3d, p2d, K, 0, rvec, tvec, false, iters, thr, conf, inliers)
 * src/keyFrameManagement.cpp:84,88.  K comes from the ctx params.  samples: NULL -> OpenCV's
 * RNG stream; else n_samples x 5 (EPNP5) or x 4 (P3P4) replay list.  inliers receives up to
 * cap ascending indices of the best model's inliers (before refinement, as OpenCV);
 * rvec/tvec are the LM-refined pose.  n == 5 / n == 4: the direct solve described at VO_PNP_*; n < 4:
 * VO_ERR_TOO_FEW_POINTS (OpenCV asserts). */
int vo_pnp_ransac(vo_ctx* ctx, const float* xyz, const float* xy, int n, int iters, double thr, double conf,
                  int min_solver, const int32_t* samples, int n_samples,
                  double rvec[3], double tvec[3], int32_t* inliers, int cap, int* n_inl);

/* Per-hypothesis view of the last vo_pnp_ransac / vo_fmat_ransac call, for parity tests:
 * PnP: models = n_h x 6 (rvec,tvec), counts = n_h inlier counts (-1 = solver failed).
 * F:   models = n_h x 3 x 9, counts = n_h x 3 (-1 = no such root).  Any pointer may be NULL. */
int vo_debug_last_pnp(vo_ctx* ctx, double* models, int32_t* counts, int cap_h, int* n_h, int* best, int* n_iters);
int vo_debug_last_fmat(vo_ctx* ctx, double* models, int32_t* counts, int cap_h, int* n_h, int* best_sample,
                       int* best_model, int* n_iters);

/* one EPnP-5 solve on the device with its intermediates (432 doubles), for parity debugging */
int vo_debug_epnp(vo_ctx* ctx, const float* obj15, const float* img10, double* dbg432);

/* ---- a-9  visualSLAM::update3dtransformation(pts, pose3x4)  src/keyFrameManagement.cpp:33-46
 * (same loop as insertKeyFrames :20-30).  M row-major 3x4 double. */
int vo_transform_points(vo_ctx* ctx, const double M[12], const float* xyz_in, int n, float* xyz_out);

/* ---- cv::cvtColor(bgr, gray, CV_BGR2GRAY) for 8-bit images (src/StereoCV.cpp:35-36; the commented-out
 * conversion of src/keyFrameManagement.cpp:53,65): bit-exact with OpenCV's fixed-point formula.  For callers
 * that hold imread's BGR frames but run the 1-channel pipeline (channels = 1).  bgr: h x w x 3, stride >= 3*w
 * bytes; gray: h x w, gray_stride >= w; both host pointers (is_device = 0) or both device pointers (1). */
int vo_bgr_to_gray(vo_ctx* ctx, const uint8_t* bgr, int stride, int is_device, uint8_t* gray, int gray_stride);

/* ---- SURVEY 8(f)-3  visualSLAM::SORcloud(ref3d, colorMap)  src/rosFuncs.cpp:9-39 (every frame at
 * src/VisualSLAM.cpp:154, on keyframes at :128): points with -z > 500 are dropped, then
 * pcl::StatisticalOutlierRemoval(meanK = mean_k, stddevMulThresh = stddev_mul; the reference uses 200 and
 * 0.01) keeps the points whose mean distance to their mean_k nearest neighbours is <= mean + mul * stddev
 * of those means.  keep_idx receives the input indices of the kept points in input order (apply it to the
 * colour vector on the caller's side); mean_dist (nullable, n floats) the per-point mean neighbour
 * distance (-1 for points that never entered the cloud).  PCL is not vendored by the reference and not
 * present in the build image: parity is against the restatement in oracle/sor.py. */
int vo_sor_cloud(vo_ctx* ctx, const float* xyz, int n, int mean_k, double stddev_mul,
                 int32_t* keep_idx, int cap, int* n_keep, float* mean_dist);

/* ---- a-11 / SURVEY 8(f)-4  dense stereo: StereoProcess::stereoMatch  src/StereoCV.cpp:21-62 and
 * StereoProcess::reprojectDisparity  src/StereoCV.cpp:221-250.
 * vo_sgbm_params mirrors the arguments of cv::StereoSGBM::create (mode = MODE_SGBM, the default the reference
 * uses); vo_sgbm_default_params fills the reference's values (1, 96, 7, 24, 96, 0, 60, 0, 3000, 5;
 * src/StereoCV.cpp:39-50).  Supported range: numDisparities a multiple of 16 up to 256, odd blockSize up to 11,
 * preFilterCap up to 126, and blockSize^2 * (2 * ftzero + 63) + P2 <= 32767 (OpenCV's int16 cost range); other
 * values, and images with width - (minDisparity + numDisparities) <= blockSize / 2 (where cv2 throws), return
 * VO_ERR_INVALID_ARG.  Image size is per call: the reference runs this on its own executable's frames. */
typedef struct vo_sgbm_params {
  int min_disparity;        /* 1    */
  int num_disparities;      /* 96   */
  int block_size;           /* 7    */
  int p1;                   /* 24   */
  int p2;                   /* 96   */
  int disp12_max_diff;      /* 0 (OpenCV treats <= 0 as 1) */
  int pre_filter_cap;       /* 60   */
  int uniqueness_ratio;     /* 0    */
  int speckle_window_size;  /* 3000 */
  int speckle_range;        /* 5    */
} vo_sgbm_params;
void vo_sgbm_default_params(vo_sgbm_params* p);
/* matcher->compute(grayL, grayR, disp): 8-bit 1-channel images (stride bytes per row) -> int16 disparity, 16x
 * fixed point, invalid = (minDisparity - 1) * 16, disp_stride bytes per row; bit-identical to cv2 4.13.0.  disp
 * may be NULL: the result then stays on the device for vo_reproject_disparity. */
int vo_sgbm_compute(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int stride, int width, int height,
                    const vo_sgbm_params* p, int16_t* disp, int disp_stride);
/* StereoProcess::stereoMatch: the same on imread's BGR frames (stride >= 3 * width), cvtColor(BGR2GRAY) on the
 * device first (src/StereoCV.cpp:35-36). */
int vo_stereo_match(vo_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, int stride, int width, int height,
                    const vo_sgbm_params* p, int16_t* disp, int disp_stride);
/* StereoProcess::reprojectDisparity: disp.convertTo(CV_32F) (the raw 16x values, as the reference does),
 * reprojectImageTo3D with Q (row-major 4x4, from the caller's stereoRectify, src/StereoCV.cpp:229), points with
 * z > 5 or z <= 0.01 skipped, y negated, raster order.  disp = NULL uses the device-resident result of the last
 * vo_sgbm_compute / vo_stereo_match of the same size.  xyz receives the points, pix_idx (nullable) their
 * y * width + x for the caller's colour lookup (lImg.at<Vec3b>(i, j), src/StereoCV.cpp:238); *n is the number of
 * points that passed (VO_ERR_CAPACITY when it exceeds cap; the first cap are written). */
int vo_reproject_disparity(vo_ctx* ctx, const int16_t* disp, int disp_stride, int width, int height, const double Q[16],
                           float* xyz, int32_t* pix_idx, int cap, int* n);
/* device milliseconds of the last SGBM call.  ms[0] = upload (+ gray conversion), ms[8] = download, *pipeline_ms
 * (nullable) = everything between.  A repeated (size, parameters) is replayed as one CUDA graph; the per-stage figures
 * ms[1..7] -- prefilter, cost volume (BT + box sums), vertical paths (the horizontal and diagonal ones run beside them
 * on two more streams), what remained of those after that, sum + winner-take-all, left-right check + median, speckle
 * filter -- exist only for calls made as plain launches, i.e. while vo_profile_enable(ctx, mask != 0) is in effect,
 * and read 0 otherwise. */
int vo_sgbm_timing(vo_ctx* ctx, float ms[9], float* pipeline_ms);
/* intermediate stages of the last SGBM call, for parity debugging: 0 = C (int16 [h][W1][D]), 1 = disparity after
 * winner-take-all, 2 = after the left-right check (both int16 [h][w]), 3 = prefilter planes (uchar4 [4][h][w]) */
int vo_debug_sgbm_stage(vo_ctx* ctx, int stage, void* out, uint64_t bytes);

/* ---- SURVEY 8(f)-2, the stages of ORB one by one (vo_orb_detect_and_compute below assembles them).  The descriptor
 * stage (orb.cpp computeOrbDescriptors): rBRIEF descriptors (32 bytes each) of caller-made keypoints on ONE 8-bit
 * pyramid level: xy = n x (x, y) in that level's pixels, angle_deg = n keypoint angles in degrees (cv::KeyPoint::angle);
 * bit-identical to cv2.ORB_create().compute(img, keypoints) of cv2 4.13.0 for octave-0 keypoints.  Keypoints must lie
 * at least 19.5 px inside the image (cv2 itself drops those closer than 31 px).  angle_deg = NULL computes the angles
 * the way detectAndCompute does (vo_orb_angles below). */
int vo_orb_describe(vo_ctx* ctx, const uint8_t* img, int stride, int width, int height, const float* xy,
                    const float* angle_deg, int n, uint8_t* desc);
/* ORB's orientation step (orb.cpp ICAngles): intensity-centroid angle in degrees of each keypoint on the UNSMOOTHED
 * level, patch radius 15, cv::fastAtan2; bit-identical to the angles cv2.ORB.detect reports for octave-0 keypoints.
 * Keypoints must lie at least 15.5 px inside the image. */
int vo_orb_angles(vo_ctx* ctx, const uint8_t* img, int stride, int width, int height, const float* xy, int n,
                  float* angle_deg);
/* ORB::create()->detectAndCompute(img, noArray(), keypoints, descriptors) as the loop detector calls it
 * (src/optimizationStuff.cpp:49-56; ORB::create() defaults with nfeatures = 500 there): 8-level INTER_LINEAR_EXACT
 * pyramid, per level FAST-9/16 (threshold 20, suppression) -> border filter (31) -> retainBest(2n) by FAST score ->
 * Harris response -> retainBest(n) -> IC_Angle -> smoothing -> rBRIEF, positions scaled back to level 0.  The keypoint
 * SET of every octave (position, response, angle) and the descriptors are bit-identical to cv2 4.13.0; the order is
 * (octave, y, x) -- OpenCV's order inside an octave is whatever std::nth_element leaves.  Both retainBest selections
 * keep every tie of their threshold (as OpenCV's do), so *n can exceed nfeatures.  All selections run on the device; the
 * call synchronises once.  octave / response / angle_deg may be NULL; *n = number of keypoints (VO_ERR_CAPACITY when it
 * exceeds cap, or 2 * nfeatures + 4096, what the library's staging holds). */
int vo_orb_detect_and_compute(vo_ctx* ctx, const uint8_t* img, int stride, int width, int height, int nfeatures, float* xy,
                              int32_t* octave, float* response, float* angle_deg, uint8_t* desc, int cap, int* n);
/* cv::FAST (TYPE_9_16, the detector ORB runs on every pyramid level with fastThreshold 20 and non-maximum
 * suppression; orb.cpp computeKeyPoints): corners in raster order with their scores (cornerScore<16>), identical to
 * cv2.FastFeatureDetector_create(threshold, nonmax).detect -- positions, order and responses.  *n = number of corners
 * found (VO_ERR_CAPACITY when it exceeds cap; the first cap are written); score may be NULL. */
int vo_fast9(vo_ctx* ctx, const uint8_t* img, int stride, int width, int height, int threshold, int nonmax_suppression,
             float* xy, float* score, int cap, int* n);
/* the response ORB ranks its keypoints by (orb.cpp HarrisResponses, blockSize 7, k = 0.04): bit-identical to
 * cv::KeyPoint::response of cv2.ORB.detect for octave-0 keypoints.  Keypoints at least 4.5 px inside the image. */
int vo_orb_harris(vo_ctx* ctx, const uint8_t* img, int stride, int width, int height, const float* xy, int n,
                  float* response);
/* the image ORB samples its descriptors from: GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) as OpenCV evaluates it on
 * a pyramid level (the float separable-filter path, not the fixed-point Gaussian) */
int vo_orb_smooth(vo_ctx* ctx, const uint8_t* img, int stride, int width, int height, uint8_t* out, int out_stride);

/* ---- a-8  Rodrigues + inversion, src/VisualSLAM.cpp:70-74,93-97: pose3x4 = [R^T | -R^T tvec]. */
int vo_pose_from_pnp(const double rvec[3], const double tvec[3], double pose3x4[12]);

/* ================= fused stage entry points = the reference's method boundaries ================= */

/* visualSLAM::denseLKtracking  src/tracking.cpp:14-28: LK + status==1 compaction.
 * ref_xy (n points) is read; ref_out/trk_out receive the m survivors in order. */
int vo_dense_lk_tracking(vo_ctx* ctx, const uint8_t* ref_img, const uint8_t* cur_img, int stride,
                         const float* ref_xy, int n, float* ref_out, float* trk_out, int* m);

/* visualSLAM::FmatThresholding  src/tracking.cpp:30-43 (thr = params.f_thr_stereo). */
int vo_fmat_thresholding(vo_ctx* ctx, const float* ref_xy, const float* trk_xy, int n,
                         float* ref_out, float* trk_out, int* m);

/* visualSLAM::stereoTriangulate  src/triangulation.cpp:73-166 (DENSE_FLAG branch):
 * grid(params.grid_step) -> LK left->right -> status compaction -> F-RANSAC(f_thr_stereo)
 * -> compaction -> triangulate with P1=K[I|0], P2=K[I|-b e1].  xyz is in the LEFT CAMERA
 * frame.  A NULL image returns VO_ERR_INVALID_ARG (the reference prints and returns). */
int vo_stereo_triangulate(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int stride,
                          float* xyz, float* xy_left, int cap, int* n);

/* visualSLAM::insertKeyFrames  src/keyFrameManagement.cpp:9-31: stereoTriangulate + transform
 * by pose3x4 (camera->world).  xyz_cam (the reference's `untransformed`) may be NULL. */
int vo_insert_keyframe(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int stride,
                       const double pose3x4[12], float* xyz_world, float* xy_left, float* xyz_cam,
                       int cap, int* n);

/* visualSLAM::PyrLKtrackFrame2Frame  src/tracking.cpp:46-91: LK ref->cur, status compaction,
 * F-RANSAC(f_thr_temporal), mask compaction (loop bound fixed to inIdx.size()).
 * Outputs: tracked 2-D, their 3-D points, and the surviving REFERENCE 2-D points
 * (inlierReferencePyrLKPts, tracking.cpp:90; may be NULL). */
int vo_track_frame(vo_ctx* ctx, const uint8_t* ref_img, const uint8_t* cur_img, int stride,
                   const float* ref_xy, const float* ref_xyz, int n,
                   float* trk_xy, float* trk_xyz, float* ref_xy_inl, int* n_out);

/* visualSLAM::PerspectiveNpointEstimation  src/keyFrameManagement.cpp:73-94:
 * vo_track_frame + solvePnPRansac(pnp_iters, pnp_thr, pnp_conf); if < pnp_min_inliers retry
 * with (pnp_retry_*); still < pnp_min_inliers -> VO_ERR_LOW_INLIERS (outputs still written).
 * attempt_used = 1 or 2. */
int vo_pnp_frame(vo_ctx* ctx, const uint8_t* ref_img, const uint8_t* cur_img, int stride,
                 const float* ref_xy, const float* ref_xyz, int n,
                 float* trk_xy, float* trk_xyz, float* ref_xy_inl, int* n_trk,
                 double rvec[3], double tvec[3], int32_t* inliers, int cap_inl, int* n_inl, int* attempt_used);

/* ================= device-resident sequence driver ================================================
 * The hot-path part of the per-frame loop of visualSLAM::initSequence (src/VisualSLAM.cpp:31,
 * 58-64,70-74,93-97,120-151) with all state (reference image pyramid, reference 2-D/3-D points)
 * kept in HBM between frames; only the pose and the counters cross PCIe per frame.  Images may
 * be host pointers (copied H2D inside the call) or device pointers (is_device != 0).
 * The stage entry points above that take images (vo_lk_track, vo_dense_lk_tracking, vo_stereo_triangulate,
 * vo_insert_keyframe, vo_track_frame, vo_pnp_frame, vo_debug_pyramid_*) build their pyramids in the same
 * slots: calling one of them ends the sequence (vo_seq_track returns VO_ERR_INVALID_ARG until the next
 * vo_seq_init). */
typedef struct vo_frame_result {
  double rvec[3], tvec[3];   /* solvePnPRansac output (world -> camera)        */
  double pose3x4[12];        /* [R|t] camera -> world, src/VisualSLAM.cpp:70-97 */
  int n_lk_in;               /* keypoints that entered the temporal LK          */
  int n_tracked;             /* after status + F-RANSAC compaction              */
  int n_inliers;             /* PnP inliers                                     */
  int attempt_used;          /* 1 or 2                                          */
  int keyframe;              /* 1 if a keyframe was inserted on this frame      */
  int n_kf_points;           /* points of the new keyframe (if keyframe)        */
  int n_lk_in_stereo;        /* keypoints that entered the stereo LK (if keyframe) */
} vo_frame_result;

int vo_seq_init(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int stride, int is_device, int* n_points);
int vo_seq_track(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int stride, int is_device,
                 int force_keyframe, vo_frame_result* out);
/* Optional: announce the NEXT frame's host images (pinned memory for a truly asynchronous copy) before
 * calling vo_seq_track for the current one.  That call enqueues their copy to device staging on a separate
 * copy stream as soon as its own first kernels are in flight; the vo_seq_track call that later passes the
 * same pointers and stride skips its own copy.  The buffers must stay unchanged until that later call has
 * returned.  Purely a transfer overlap: results are the same with or without it. */
int vo_seq_prefetch(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int stride);
/* The same announcement for either kind of pointer: is_device = 0 is vo_seq_prefetch; is_device != 0 announces a
 * frame that is already resident in device memory (nothing is copied).  When the keyframe policy inserts a keyframe
 * on every frame, an announced frame (of either kind) lets the library build the NEXT frame's left image pyramid -- which
 * depends on nothing but the image -- on a third stream while the current frame is processed and enqueue that frame's
 * tracking LK at the end of the current call, behind the epilogue that produces its input points (default), or, opt-in
 * (VO_B200_LOOKAHEAD=1 with VO_B200_SEQ_HOST=1), run the next frame's whole temporal LK there.  The announced image
 * must stay unchanged until the vo_seq_track call that passes it has returned; the next vo_seq_track call must pass
 * the announced pointers to use the work done ahead, any other image simply discards it.
 * Results are identical with and without the announcement. */
int vo_seq_announce(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int stride, int is_device);
/* current reference set (after the last vo_seq_* call); any pointer may be NULL */
int vo_seq_get_reference(vo_ctx* ctx, float* xy, float* xyz, int cap, int* n);

/* ================= harness / measurement helpers ================================================= */
void* vo_cuda_stream(vo_ctx* ctx);   /* cudaStream_t every kernel of this ctx is launched on */
int vo_sync(vo_ctx* ctx);
/* per-kernel device timing: `mask` is a bit set of kernel families (bit k = VO_K_*; 0 = off,
 * -1 = all).  Every launch of a selected family is bracketed by CUDA events on the ctx stream;
 * vo_profile_read returns launches and summed milliseconds. */
enum { VO_K_PYRAMID = 0, VO_K_LK = 1, VO_K_COMPACT = 2, VO_K_FMAT_SOLVE = 3, VO_K_FMAT_SCORE = 4,
       VO_K_TRIANGULATE = 5, VO_K_PNP_SOLVE = 6, VO_K_PNP_SCORE = 7, VO_K_PNP_REFINE = 8,
       VO_K_SELECT = 9, VO_K_MISC = 10, VO_K_COUNT = 11 };
int vo_profile_enable(vo_ctx* ctx, int mask);
int vo_profile_read(vo_ctx* ctx, int kernel, int64_t* launches, double* ms, int reset);
/* timeline of the bracketed launches since the last vo_profile_read: rows of 4 floats
 * (chain 0/1, kernel family, start ms relative to the first launch, duration ms) */
int vo_debug_timeline(vo_ctx* ctx, float* rows, int cap, int* n);
int64_t vo_launch_count(vo_ctx* ctx);          /* kernels launched by this ctx since creation */
/* Cumulative LK work counters since vo_create: (point, level) pairs processed and LK iterations
 * executed -- the units of the LK roofline in DESIGN.md.  Take differences around a region. */
int vo_lk_work(vo_ctx* ctx, int64_t* point_levels, int64_t* iterations);
/* Of those, how many took the sequential float-chain path (a partial window sum could reach 2^24, where
 * OpenCV's float accumulators start rounding): (point, level) window extractions and iterations. */
int vo_lk_slow_paths(vo_ctx* ctx, int64_t* window_sums, int64_t* iterations);
/* FP32 issue-rate microbenchmark (dependent FFMA chains on every SM): TFLOP/s achieved. */
int vo_measure_fp32_peak(vo_ctx* ctx, double* tflops);
/* INT32 issue-rate microbenchmark (dependent IMAD chains on every SM): Tops/s achieved (multiply-add = 2 ops);
 * the denominator of the LK roofline (integer kernel). */
int vo_measure_int32_peak(vo_ctx* ctx, double* tops);
/* synthetic scene renderer (harness only; same scene as oracle/synth.py): renders frame
 * `frame` of scene `seed` into a DEVICE buffer of width*height bytes.  eye 0 = left, 1 = right. */
int vo_synth_render_dev(vo_ctx* ctx, int seed, int frame, int eye, uint8_t* out_dev);
/* pinned host memory + plain device memory for the harness */
int vo_alloc_host(void** p, uint64_t bytes);
int vo_free_host(void* p);
int vo_alloc_dev(vo_ctx* ctx, void** p, uint64_t bytes);
int vo_free_dev(vo_ctx* ctx, void* p);
int vo_memcpy_d2h(vo_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);
int vo_memcpy_h2d(vo_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VO_B200_H */
