// Host build of ros_stereo_slam_b200/csrc/cvmath.cuh for CPU-side unit tests.
// TEST TOOL ONLY: it lets tests/test_hostmath.py pin the device numerics against cv2
// in a container without a GPU.  libvo_b200.so never links or calls this.
#include "../../ros_stereo_slam_b200/csrc/cvmath.cuh"
#include "../../ros_stereo_slam_b200/csrc/fmat7.cuh"
#include <string.h>
using namespace vo;
extern "C" {
void hm_svd(const double* A, int n, double* w, double* u, double* vt) {
  switch (n) {
    case 3: svd_square<3>(A, w, u, vt); break;
    case 4: svd_square<4>(A, w, u, vt); break;
    case 12: svd_square<12>(A, w, u, vt); break;
  }
}
void hm_solve(const double* A, const double* b, int n, double* x) {
  if (n == 3) solve_svd<6, 3>(A, b, x);
  if (n == 4) solve_svd<6, 4>(A, b, x);
  if (n == 5) solve_svd<6, 5>(A, b, x);
}
// cv::solve(A 6 x n, DECOMP_SVD) through the runtime-size Jacobi pieces; wavefront != 0 walks the pairs in the order and
// with the completion rule of the device's lane-group schedule (jacobi_warp.cuh, jacobi_sweeps_groups): pair (i, j) of
// sweep s at step s * n + i + j, two sweeps in flight, the sweep after an unchanged one started speculatively
void hm_solve_rt(const double* A, const double* b, int n, int wavefront, double* x) {
  double at[30], w[5], v[25], W[8];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 6; j++) at[i * 6 + j] = A[j * n + i];
  if (!wavefront) {
    jacobi_svd_rt<6, 5>(at, w, v, n);
  } else {
    const int max_iter = 30, H = n >> 1, P = n, D = 2 * n - 3;
    jacobi_rt_init<6, 5>(at, W, v, n);
    bool done = n < 2, chg_old = false, chg_new = false;
    int s_new = 0;
    for (int step = 1; !done; step++) {
      const int t_new = step - s_new * P;
      bool rot[8] = {false, false, false, false, false, false, false, false};
      for (int idx = 0; idx < 2 * H; idx++) {      // the lanes of the group (their pairs are disjoint)
        const bool is_new = idx >= H;
        const int sweep = is_new ? s_new : s_new - 1;
        const int t = is_new ? t_new : t_new + P;
        if (sweep >= 0 && sweep < max_iter && t <= D) {
          const int li = is_new ? idx - H : idx;
          const int i0 = t - (n - 1) > 0 ? t - (n - 1) : 0;
          const int i = i0 + li, j = t - i;
          if (i < j) rot[idx] = jacobi_rt_rotate<6, 5>(at, W, v, i, j);
        }
      }
      for (int idx = 0; idx < H; idx++) chg_old |= rot[idx];
      for (int idx = H; idx < 2 * H; idx++) chg_new |= rot[idx];
      if (D > P) {
        if (s_new >= 1 && t_new + P == D && (!chg_old || s_new - 1 == max_iter - 1)) done = true;
      } else {
        if (t_new == D && (!chg_new || s_new == max_iter - 1)) done = true;
      }
      if (t_new == P) {
        chg_old = chg_new;
        chg_new = false;
        s_new++;
      }
    }
    jacobi_rt_finish<6, 5>(at, w, v, n);
  }
  svd_backsubst_6xn(at, w, v, b, x, n);
}
void hm_invert3(const double* A, double* inv) { invert3_svd(A, inv); }
void hm_mtm12(const double* M, int rows, int fma_, double* out) {
  if (fma_) mul_transposed<12, true>(M, rows, out); else mul_transposed<12, false>(M, rows, out);
}
void hm_epnp5(const float* obj, const float* img, const double* K4, int fma_, double* rvec, double* tvec, double* R) {
  Intrinsics K{K4[0], K4[1], K4[2], K4[3]};
  double t[3];
  if (fma_) epnp5<true>(obj, img, K, R, t); else epnp5<false>(obj, img, K, R, t);
  rodrigues_mat2vec(R, rvec);
  for (int i = 0; i < 3; i++) tvec[i] = t[i];
}
void hm_rodrigues_v2m(const double* r, double* R) { rodrigues_vec2mat(r, R); }
void hm_triangulate(const double* P1, const double* P2, const float* xy1, const float* xy2, int n, float* xyz, float* h4) {
  for (int i = 0; i < n; i++)
    triangulate_dlt(P1, P2, xy1[2 * i], xy1[2 * i + 1], xy2[2 * i], xy2[2 * i + 1], xyz + 3 * i, h4 ? h4 + 4 * i : nullptr);
}
void hm_epnp5_dbg(const float* obj, const float* img, const double* K4, double* dbg) {
  Intrinsics K{K4[0], K4[1], K4[2], K4[3]};
  double R[9], t[3];
  epnp5<false>(obj, img, K, R, t, dbg);
  for (int i = 0; i < 9; i++) dbg[420 + i] = R[i];
  for (int i = 0; i < 3; i++) dbg[429 + i] = t[i];
}
int hm_fmat7(const float* m1, const float* m2, double* F) { return fmat_7point(m1, m2, F); }
int hm_solve_cubic(const double* c, double* r) { return solve_cubic(c, r); }
}
