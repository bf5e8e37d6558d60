// mirror_main.cpp -- test driver for the C++ mirror of the reference interface (include/vo_b200.hpp).
// Reads raw frames written by tests/test_cpp_mirror.py, runs the reference's own call sequence
//   stereoTriangulate(L0, R0) -> PerspectiveNpointEstimation(L0, L1, ...) -> insertKeyFrames(L1, R1, pose)
// through vo::visualSLAM and dumps the results as raw binary for the Python side to compare with the
// ctypes path and the oracle.  Usage: mirror_main <dir> <width> <height> <channels>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "vo_b200.hpp"

static std::vector<uint8_t> read_file(const std::string& path, size_t bytes) {
  std::vector<uint8_t> buf(bytes);
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f || std::fread(buf.data(), 1, bytes, f) != bytes) {
    std::fprintf(stderr, "cannot read %s\n", path.c_str());
    std::exit(2);
  }
  std::fclose(f);
  return buf;
}

template <class T>
static void write_vec(const std::string& path, const std::vector<T>& v) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) std::exit(3);
  if (!v.empty()) std::fwrite(v.data(), sizeof(T), v.size(), f);
  std::fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 5) return 1;
  const std::string dir = argv[1];
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), cn = std::atoi(argv[4]);
  const size_t bytes = (size_t)w * h * cn;
  const auto L0 = read_file(dir + "/L0.raw", bytes), R0 = read_file(dir + "/R0.raw", bytes);
  const auto L1 = read_file(dir + "/L1.raw", bytes), R1 = read_file(dir + "/R1.raw", bytes);
  auto view = [&](const std::vector<uint8_t>& b) { return vo::Image{b.data(), h, w, w * cn}; };

  vo_params p;
  vo_default_params(&p);
  p.width = w;
  p.height = h;
  p.channels = cn;
  try {
    vo::visualSLAM slam(&p);
    std::vector<vo::KeyPoint> grid = slam.denseKeypointExtractor(view(L0), p.grid_step);
    std::vector<vo::Point3f> ref3d;
    std::vector<vo::Point2f> ref2d;
    slam.stereoTriangulate(view(L0), view(R0), ref3d, ref2d);           // src/VisualSLAM.cpp:31
    std::vector<vo::Point2f> trk2d;
    std::vector<vo::Point3f> trk3d;
    double rvec[3], tvec[3];
    std::vector<int> inliers;
    slam.PerspectiveNpointEstimation(view(L0), view(L1), ref2d, ref3d, trk2d, trk3d, rvec, tvec, inliers);   // :64
    double pose[12];
    vo_pose_from_pnp(rvec, tvec, pose);                                                                      // :70-97
    std::vector<vo::Point2f> kf2d;
    std::vector<vo::Point3f> kf3d;
    slam.insertKeyFrames(0, view(L1), view(R1), pose, kf2d, kf3d);                                         // :123
    std::vector<vo::Point3f> moved = slam.update3dtransformation(slam.untransformed, pose);

    write_vec(dir + "/ref3d.bin", ref3d);
    write_vec(dir + "/ref2d.bin", ref2d);
    write_vec(dir + "/trk2d.bin", trk2d);
    write_vec(dir + "/trk3d.bin", trk3d);
    write_vec(dir + "/inliers.bin", inliers);
    write_vec(dir + "/pose.bin", std::vector<double>{rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]});
    write_vec(dir + "/kf2d.bin", kf2d);
    write_vec(dir + "/kf3d.bin", kf3d);
    write_vec(dir + "/moved.bin", moved);
    std::printf("grid %zu stereo %zu tracked %zu inliers %zu keyframe %zu shutdown %d\n", grid.size(), ref2d.size(),
                trk2d.size(), inliers.size(), kf2d.size(), (int)slam.SHUTDOWN_FLAG);
    if (cn == 1) {
      // the loop detector's per-frame feature extraction (reference src/optimizationStuff.cpp:49-56)
      vo::ORB orb(slam.ctx());
      std::vector<vo::OrbKeyPoint> kps;
      std::vector<uint8_t> desc;
      orb.detectAndCompute(view(L1), kps, desc);
      write_vec(dir + "/orb_kps.bin", kps);
      write_vec(dir + "/orb_desc.bin", desc);
      std::printf("orb %zu keypoints\n", kps.size());
    }
    if (cn == 3) {
      // the dense-stereo executable's loop body (reference src/StereoCV.cpp:254-259): stereoMatch -> reprojectDisparity
      vo::StereoProcess sp(slam.ctx());
      const vo::BgrImage bl{L0.data(), h, w, 3 * w}, br{R0.data(), h, w, 3 * w};
      vo::Disparity disp = sp.stereoMatch(bl, br);
      std::vector<vo::Point3f> cloud, colors;
      sp.reprojectDisparity(disp, cloud, colors);            // t = +baseline: the reference's Q, no point passes
      const size_t n_ref = cloud.size();
      sp.baseline = -sp.baseline;
      sp.reprojectDisparity(disp, cloud, colors);
      write_vec(dir + "/disp.bin", disp.data);
      write_vec(dir + "/cloud.bin", cloud);
      write_vec(dir + "/colors.bin", colors);
      std::printf("sgbm cloud %zu (reference Q: %zu)\n", cloud.size(), n_ref);
    }
  } catch (const vo::Error& e) {
    std::fprintf(stderr, "vo::Error %d: %s\n", e.code, e.what());
    return 4;
  }
  return 0;
}
