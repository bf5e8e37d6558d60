// vo_b200.hpp -- header-only C++ mirror of the reference's hot-path member functions over the
// C ABI (vo_b200.h).  Same names, argument order and in/out conventions as the members of
// `class visualSLAM` declared at reference include/visualSLAM.h:152-169, with plain structs
// standing in for cv::Point2f / cv::Point3f / cv::KeyPoint and an 8-bit image view standing in
// for cv::Mat, so that the ROS node can swap the bodies of those members for one-line calls
// (INTEGRATION.md shows the cv::Mat / std::vector adaptors).
//
// Error behaviour: the reference prints and continues (NULL image), or sets SHUTDOWN_FLAG
// (low inliers); this mirror does the same.  Any other failure of the library throws
// vo::Error -- there is no CPU fallback to hide it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "vo_b200.h"

namespace vo {

struct Point2f { float x, y; };
struct Point3f { float x, y, z; };
struct KeyPoint { Point2f pt; float size; float response; };
struct Image {               // view of an 8-bit single-channel image (cv::Mat::data / step)
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0, step = 0;
};

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

class visualSLAM {
 public:
  bool SHUTDOWN_FLAG = false;                         // reference include/visualSLAM.h:72
  std::vector<Point2f> refDrawPts, trackedDrawPts;    // side effects of PyrLKtrackFrame2Frame
  std::vector<Point2f> inlierReferencePyrLKPts;       // (reference src/tracking.cpp:88-90)
  std::vector<Point3f> untransformed;                 // side effect of insertKeyFrames (:18)
  int lastPnPAttempt = 1;

  explicit visualSLAM(const vo_params* params = nullptr) {
    if (params) p_ = *params; else vo_default_params(&p_);
    check(vo_create(&p_, &ctx_));
  }
  ~visualSLAM() { vo_destroy(ctx_); }
  visualSLAM(const visualSLAM&) = delete;
  visualSLAM& operator=(const visualSLAM&) = delete;
  vo_ctx* ctx() { return ctx_; }

  // reference src/tracking.cpp:4-12
  std::vector<KeyPoint> denseKeypointExtractor(const Image& img, int stepSize) {
    std::vector<float> xy(2 * (size_t)p_.max_points);
    int n = 0;
    check(vo_grid_keypoints(ctx_, img.rows, img.cols, stepSize, xy.data(), p_.max_points, &n));
    std::vector<KeyPoint> out(n);
    for (int i = 0; i < n; i++) out[i] = KeyPoint{{xy[2 * i], xy[2 * i + 1]}, (float)stepSize, 0.f};
    return out;
  }

  // reference src/tracking.cpp:14-28: both vectors are replaced by the status==1 subsets
  void denseLKtracking(const Image& refImg, const Image& curImg, std::vector<Point2f>& refPts,
                       std::vector<Point2f>& trackPts) {
    const int n = (int)refPts.size();
    std::vector<Point2f> r(n), t(n);
    int m = 0;
    check(vo_dense_lk_tracking(ctx_, refImg.data, curImg.data, refImg.step, f(refPts), n, f(r), f(t), &m));
    r.resize(m); t.resize(m);
    refPts.swap(r); trackPts.swap(t);
  }

  // reference src/tracking.cpp:30-43
  void FmatThresholding(std::vector<Point2f>& refPts, std::vector<Point2f>& trkPts) {
    const int n = (int)refPts.size();
    std::vector<Point2f> r(n), t(n);
    int m = 0;
    check(vo_fmat_thresholding(ctx_, f(refPts), f(trkPts), n, f(r), f(t), &m));
    r.resize(m); t.resize(m);
    refPts.swap(r); trkPts.swap(t);
  }

  // reference src/triangulation.cpp:73-166: outputs overwritten; NULL image prints and returns
  void stereoTriangulate(const Image& im1, const Image& im2, std::vector<Point3f>& ref3dPts,
                         std::vector<Point2f>& ref2dPts) {
    if (!im1.data || !im2.data) { std::printf("NULL IMG\n"); return; }
    std::vector<Point3f> xyz(p_.max_points);
    std::vector<Point2f> xy(p_.max_points);
    int n = 0;
    check(vo_stereo_triangulate(ctx_, im1.data, im2.data, im1.step, f(xyz), f(xy), p_.max_points, &n));
    xyz.resize(n); xy.resize(n);
    ref3dPts.swap(xyz); ref2dPts.swap(xy);
  }

  // reference src/tracking.cpp:46-91: outputs are appended to (the caller passes empty vectors)
  void PyrLKtrackFrame2Frame(const Image& refimg, const Image& curImg, std::vector<Point2f> refPts,
                             std::vector<Point3f> ref3dpts, std::vector<Point2f>& refRetpts,
                             std::vector<Point3f>& ref3dretPts) {
    const int n = (int)refPts.size();
    std::vector<Point2f> t2(n), r2(n);
    std::vector<Point3f> t3(n);
    int m = 0;
    check(vo_track_frame(ctx_, refimg.data, curImg.data, refimg.step, f(refPts), f(ref3dpts), n, f(t2), f(t3), f(r2), &m));
    t2.resize(m); t3.resize(m); r2.resize(m);
    refRetpts.insert(refRetpts.end(), t2.begin(), t2.end());
    ref3dretPts.insert(ref3dretPts.end(), t3.begin(), t3.end());
    refDrawPts = r2; trackedDrawPts = refRetpts; inlierReferencePyrLKPts = r2;
  }

  // reference src/keyFrameManagement.cpp:9-31; pose4dTransform is the 3x4 CV_64F [R|t]
  void insertKeyFrames(int /*start*/, const Image& imL, const Image& imR, const double pose4dTransform[12],
                       std::vector<Point2f>& ftrPts, std::vector<Point3f>& ref3dCoords) {
    ftrPts.clear(); ref3dCoords.clear();
    if (!imL.data || !imR.data) { std::printf("NULL IMG\n"); return; }
    std::vector<Point3f> w(p_.max_points), cam(p_.max_points);
    std::vector<Point2f> xy(p_.max_points);
    int n = 0;
    check(vo_insert_keyframe(ctx_, imL.data, imR.data, imL.step, pose4dTransform, f(w), f(xy), f(cam), p_.max_points, &n));
    w.resize(n); cam.resize(n); xy.resize(n);
    untransformed.swap(cam); ref3dCoords.swap(w); ftrPts.swap(xy);
  }

  // reference src/keyFrameManagement.cpp:33-46
  std::vector<Point3f> update3dtransformation(std::vector<Point3f>& pt3d, const double pose4dTransform[12]) {
    std::vector<Point3f> out(pt3d.size());
    check(vo_transform_points(ctx_, pose4dTransform, f(pt3d), (int)pt3d.size(), f(out)));
    return out;
  }

  // reference src/keyFrameManagement.cpp:73-94.  The reference ignores its prevImg/curImg arguments and
  // reads the members referenceImg/currentImage; here they are the images that are used.
  void PerspectiveNpointEstimation(const Image& prevImg, const Image& curImg, std::vector<Point2f>& ref2dPoints,
                                   std::vector<Point3f>& ref3dPoints, std::vector<Point2f>& tracked2dPoints,
                                   std::vector<Point3f>& tracked3dPoints, double rvec[3], double tvec[3],
                                   std::vector<int>& inliers) {
    const int n = (int)ref2dPoints.size();
    std::vector<Point2f> t2(n), r2(n);
    std::vector<Point3f> t3(n);
    std::vector<int32_t> inl(n > 0 ? n : 1);
    int m = 0, ni = 0, att = 1;
    const int r = vo_pnp_frame(ctx_, prevImg.data, curImg.data, prevImg.step, f(ref2dPoints), f(ref3dPoints), n, f(t2),
                               f(t3), f(r2), &m, rvec, tvec, inl.data(), (int)inl.size(), &ni, &att);
    if (r != VO_OK && r != VO_ERR_LOW_INLIERS) check(r);
    t2.resize(m); t3.resize(m); r2.resize(m);
    tracked2dPoints.insert(tracked2dPoints.end(), t2.begin(), t2.end());
    tracked3dPoints.insert(tracked3dPoints.end(), t3.begin(), t3.end());
    refDrawPts = r2; trackedDrawPts = tracked2dPoints; inlierReferencePyrLKPts = r2;
    inliers.assign(inl.begin(), inl.begin() + ni);
    lastPnPAttempt = att;
    if (r == VO_ERR_LOW_INLIERS) {
      std::fprintf(stderr, "low inlier count (%d) after the relaxed retry\n", ni);
      SHUTDOWN_FLAG = true;   // reference src/keyFrameManagement.cpp:89-92
    }
  }

 private:
  static float* f(std::vector<Point2f>& v) { return reinterpret_cast<float*>(v.data()); }
  static float* f(std::vector<Point3f>& v) { return reinterpret_cast<float*>(v.data()); }
  static void check(int r) {
    if (r != VO_OK) throw Error(r, std::string(vo_strerror(r)) + ": " + vo_last_error());
  }
  vo_params p_;
  vo_ctx* ctx_ = nullptr;
};

// Mirror of the dense-stereo members of the reference's `class StereoProcess` (reference include/stereoCV.h:63-64,
// src/StereoCV.cpp:21-62,221-250).  The reference reads the frames itself (getImg + imread); here the caller hands
// over the BGR frames it read, everything after imread runs on the GPU.
struct BgrImage {            // view of an 8-bit 3-channel image as imread returns it
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0, step = 0;
};
struct Disparity {           // CV_16S disparity, 16x fixed point (what StereoSGBM::compute returns)
  std::vector<int16_t> data;
  int rows = 0, cols = 0;
};

class StereoProcess {
 public:
  double baseline = 0.5707;                                   // reference include/stereoCV.h:39-43
  double focal_x = 7.188560000000e+02, cx = 6.071928000000e+02;
  double focal_y = 7.188560000000e+02, cy = 1.852157000000e+02;
  vo_sgbm_params sgbm;                                        // StereoSGBM::create arguments, src/StereoCV.cpp:39-50

  explicit StereoProcess(vo_ctx* ctx) : ctx_(ctx) { vo_sgbm_default_params(&sgbm); }

  // reference src/StereoCV.cpp:21-62: cvtColor(BGR2GRAY) x 2 + matcher->compute
  Disparity stereoMatch(const BgrImage& im1, const BgrImage& im2) {
    lImg_ = im1;
    Disparity d;
    d.rows = im1.rows; d.cols = im1.cols;
    d.data.resize((size_t)im1.rows * im1.cols);
    check(vo_stereo_match(ctx_, im1.data, im2.data, im1.step, im1.cols, im1.rows, &sgbm, d.data.data(), 2 * im1.cols));
    return d;
  }

  // reference src/StereoCV.cpp:221-250.  Q is what stereoRectify(K, 0, K, 0, size, I, (baseline, 0, 0)) returns
  // there (:224-229); for those arguments (no rotation, equal intrinsics) it has the closed form below -- OpenCV
  // 4.13.0 rounds the principal point through float, which is reproduced (bit-equal to cv2.stereoRectify).  With
  // t = +baseline every reprojected z is negative and the gate of :240 drops every point: the reference's behaviour.
  void reprojectDisparity(const Disparity& disp, std::vector<Point3f>& reproject3dPoints, std::vector<Point3f>& colorMap) {
    reproject3dPoints.clear(); colorMap.clear();
    const double Q[16] = {1, 0, 0, -(double)(float)cx, 0, 1, 0, -(double)(float)cy, 0, 0, 0, focal_x,
                          0, 0, -1.0 / baseline, 0};
    const int n = disp.rows * disp.cols;
    std::vector<Point3f> pts(n);
    std::vector<int32_t> idx(n);
    int m = 0;
    check(vo_reproject_disparity(ctx_, disp.data.data(), 2 * disp.cols, disp.cols, disp.rows, Q,
                                 reinterpret_cast<float*>(pts.data()), idx.data(), n, &m));
    pts.resize(m);
    reproject3dPoints.swap(pts);
    colorMap.resize(m);
    for (int k = 0; k < m; k++) {       // Vec3b colors = lImg.at<Vec3b>(i, j), src/StereoCV.cpp:238,245
      const int i = idx[k] / disp.cols, j = idx[k] % disp.cols;
      const uint8_t* c = lImg_.data ? lImg_.data + (size_t)i * lImg_.step + 3 * j : nullptr;
      colorMap[k] = c ? Point3f{(float)c[0], (float)c[1], (float)c[2]} : Point3f{0, 0, 0};
    }
  }

 private:
  static void check(int r) {
    if (r != VO_OK) throw Error(r, std::string(vo_strerror(r)) + ": " + vo_last_error());
  }
  vo_ctx* ctx_;
  BgrImage lImg_;
};

// Mirror of the loop detector's feature extraction (reference src/optimizationStuff.cpp:49-56:
// `Ptr<ORB> orb = ORB::create(); orb->detectAndCompute(img, noArray(), kps, descriptors)`).  Keypoints carry what
// cv::KeyPoint carries; descriptors are n x 32 bytes, the rows DBoW2's FORB takes.
struct OrbKeyPoint { Point2f pt; float size, angle, response; int octave; };

class ORB {
 public:
  explicit ORB(vo_ctx* ctx, int nfeatures = 500) : ctx_(ctx), nfeatures_(nfeatures) {}

  void detectAndCompute(const Image& img, std::vector<OrbKeyPoint>& keypoints, std::vector<uint8_t>& descriptors) {
    const int cap = 2 * nfeatures_ + 4096;
    std::vector<float> xy(2 * (size_t)cap), resp(cap), ang(cap);
    std::vector<int32_t> oct(cap);
    descriptors.assign((size_t)cap * 32, 0);
    int n = 0;
    const int r = vo_orb_detect_and_compute(ctx_, img.data, img.step, img.cols, img.rows, nfeatures_, xy.data(), oct.data(),
                                            resp.data(), ang.data(), descriptors.data(), cap, &n);
    if (r != VO_OK) throw Error(r, std::string(vo_strerror(r)) + ": " + vo_last_error());
    descriptors.resize((size_t)n * 32);
    keypoints.resize(n);
    float scale[8];
    for (int l = 0; l < 8; l++) scale[l] = (float)std::pow((double)1.2f, (double)l);     // orb.cpp getScale
    for (int i = 0; i < n; i++)
      keypoints[i] = OrbKeyPoint{{xy[2 * i], xy[2 * i + 1]}, 31.f * scale[oct[i]], ang[i], resp[i], oct[i]};
  }

 private:
  vo_ctx* ctx_;
  int nfeatures_;
};

}  // namespace vo
