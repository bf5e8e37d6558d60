// refine.cu -- K8: Levenberg-Marquardt pose refinement on the RANSAC inliers, i.e. the final
// solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess = best RANSAC model) inside
// cv::solvePnPRansac (reference call site src/keyFrameManagement.cpp:84,88).
//
// Restates OpenCV's CvLevMarq state machine (6 parameters, criteria 20 iterations /
// FLT_EPSILON relative step, lambda = 10^k starting at k=-3, damped normal equations solved
// by SVD) over the pixel reprojection error of cvProjectPoints2 with analytic Jacobians.
// One thread-block CLUSTER (8 CTAs): the residual/Jacobian pass is data-parallel over the
// inliers with an FP64 tree reduction that crosses CTAs through distributed shared memory;
// the 6x6 solve and the state machine run on thread 0 of every CTA.  Parallel summation
// reorders OpenCV's sequential sums (relative 1e-13), far inside the 1e-4 rad / 1e-3 m
// pose tolerance.
#include <cooperative_groups.h>

#include "common.cuh"
#include "cvmath.cuh"

namespace vo {

constexpr int REF_THREADS = 512;
constexpr int REF_CLUSTER = 8;   // CTAs per thread-block cluster (portable maximum)
constexpr int NACC = 28;         // 21 (JtJ upper) + 6 (JtErr) + 1 (|err|^2)
constexpr int REF_RPT = 4;       // inliers per thread kept in registers across the LM passes (covers 16,384 inliers)

// CvLevMarq's damping factors 10^k as OpenCV forms them, exp(k * log(10.)), k = -16 .. 16: evaluated once on the host
// instead of two libm calls on the one thread every CTA waits for
__constant__ double c_lm_lambda[33];

int refine_init() {
  static int rc = [] {
    double tab[33];
    for (int k = -16; k <= 16; k++) tab[k + 16] = exp(k * log(10.));
    return cudaMemcpyToSymbol(c_lm_lambda, tab, sizeof(tab)) == cudaSuccess ? VO_OK : VO_ERR_CUDA;
  }();
  return rc;
}

struct PoseJac {
  double R[9];
  double dRdr[27];  // dR_k/dr_i at [i*9+k]
};

__device__ void rodrigues_with_jacobian(const double r_[3], PoseJac& o) {
  double rx = r_[0], ry = r_[1], rz = r_[2];
  const double theta = sqrt(rx * rx + ry * ry + rz * rz);
  if (theta < DBL_EPSILON) {
    for (int i = 0; i < 9; i++) o.R[i] = (i % 4 == 0) ? 1. : 0.;
    for (int i = 0; i < 27; i++) o.dRdr[i] = 0;
    o.dRdr[5] = o.dRdr[15] = o.dRdr[19] = -1;
    o.dRdr[7] = o.dRdr[11] = o.dRdr[21] = 1;
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
  rx *= itheta; ry *= itheta; rz *= itheta;
  const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
  const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int k = 0; k < 9; k++) o.R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
  const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0,
                           0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
  const double d_r_x_[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
  for (int i = 0; i < 3; i++) {
    const double ri = i == 0 ? rx : i == 1 ? ry : rz;
    const double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
    const double a3 = (c - s * itheta) * ri, a4 = s * itheta;
    for (int k = 0; k < 9; k++)
      o.dRdr[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x_[i * 9 + k];
  }
}

// 6x6 linear solve (damped normal equations), Gaussian elimination with partial pivoting.
// OpenCV solves the same system with DECOMP_SVD; for this well-conditioned SPD system the two
// agree to ~1e-13 relative, far inside the pose tolerance, and this is ~20x shorter.
__device__ void solve6(const double* A_in, const double* b_in, double* x) {
  double A[36], b[6];
  for (int i = 0; i < 36; i++) A[i] = A_in[i];
  for (int i = 0; i < 6; i++) b[i] = b_in[i];
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double mx = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < 6; i++)
      if (fabs(A[i * 6 + k]) > mx) {
        mx = fabs(A[i * 6 + k]);
        piv = i;
      }
    if (mx == 0) {
      for (int i = 0; i < 6; i++) x[i] = 0;
      return;
    }
    if (piv != k) {
      for (int j = 0; j < 6; j++) {
        const double t = A[k * 6 + j];
        A[k * 6 + j] = A[piv * 6 + j];
        A[piv * 6 + j] = t;
      }
      const double t = b[k];
      b[k] = b[piv];
      b[piv] = t;
    }
    const double inv = 1. / A[k * 6 + k];
    for (int i = k + 1; i < 6; i++) {
      const double f = A[i * 6 + k] * inv;
      for (int j = k; j < 6; j++) A[i * 6 + j] -= f * A[k * 6 + j];
      b[i] -= f * b[k];
    }
  }
  for (int i = 5; i >= 0; i--) {
    double sum = b[i];
    for (int j = i + 1; j < 6; j++) sum -= A[i * 6 + j] * x[j];
    x[i] = sum / A[i * 6 + i];
  }
}

// One thread-block cluster of REF_CLUSTER CTAs.  Every CTA accumulates its slice of the inliers,
// reduces to 28 partial sums in its own shared memory, and after a cluster barrier every CTA
// adds up all partials through distributed shared memory in the same order -- so all CTAs hold
// bit-identical sums and run the (cheap) LM state machine redundantly instead of broadcasting.
__global__ void __cluster_dims__(REF_CLUSTER, 1, 1) __launch_bounds__(REF_THREADS, 1)
pnp_refine_kernel(const float3* __restrict__ xyz, const float2* __restrict__ xy, const int32_t* __restrict__ idx,
                  const int* __restrict__ n_inl_p, const double* __restrict__ models, const int* __restrict__ sel,
                  Intrinsics K, double* __restrict__ pose_out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s_warp[(REF_THREADS / 32) * NACC];
  __shared__ double s_part[2][NACC];   // this CTA's partial sums (double-buffered across evaluations)
  __shared__ double s_sum[NACC];
  __shared__ PoseJac s_pj;
  __shared__ double s_param[6];
  __shared__ int s_state;  // 0 = need J+err at s_param, 1 = need err only, 2 = done
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned rank = cluster.block_rank();
  const int n = *n_inl_p;
  const int best = sel[0];
  if (best < 0 || n < 3) {   // uniform over the cluster
    if (t == 0 && rank == 0) pose_out[6] = -1;
    return;
  }
  // thread-0 state (CvLevMarq), replicated in every CTA; in shared memory so that the other 511 threads do not carry
  // 100 registers of it
  __shared__ double param[6], prev_param[6], JtJ[36], JtErr[6];
  double prev_err_norm = DBL_MAX, err_norm = 0;
  int lambda_lg10 = -3, iters = 0;
  const int max_iter = 20;
  const double epsilon = (double)FLT_EPSILON;
  if (t == 0) {
    const double* m = models + (size_t)best * 16;
    for (int i = 0; i < 6; i++) {
      param[i] = m[i];
      prev_param[i] = m[i];
      s_param[i] = m[i];
    }
    s_state = 0;
  }
  __syncthreads();

  // this thread's inliers: gathered once (two dependent global loads each), reused by every pass
  const int first = rank * REF_THREADS + t;
  float3 Pk[REF_RPT];
  float2 qk[REF_RPT];
#pragma unroll
  for (int k = 0; k < REF_RPT; k++) {
    const int i = first + k * REF_CLUSTER * REF_THREADS;
    Pk[k] = make_float3(0, 0, 1);
    qk[k] = make_float2(0, 0);
    if (i < n) {
      const int id = idx[i];
      Pk[k] = xyz[id];
      qk[k] = xy[id];
    }
  }

  int buf = 0;
  for (int guard = 0; guard < 1000; guard++) {
    const int state = s_state;
    if (state == 2) break;
    if (t == 0) rodrigues_with_jacobian(s_param, s_pj);
    __syncthreads();
    // J is accumulated on every pass: when a candidate is accepted, the next LM step needs J at
    // exactly this parameter vector, and recomputing it would cost a second pass over the inliers
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = 0;
    const double* R = s_pj.R;
    const double* dR = s_pj.dRdr;
    const double t0 = s_param[3], t1 = s_param[4], t2 = s_param[5];
    auto accumulate = [&](const float3 P, const float2 q) {
      const double X = P.x, Y = P.y, Z = P.z;
      double x = R[0] * X + R[1] * Y + R[2] * Z + t0;
      double y = R[3] * X + R[4] * Y + R[5] * Z + t1;
      double z = R[6] * X + R[7] * Y + R[8] * Z + t2;
      z = z ? 1. / z : 1;
      x *= z;
      y *= z;
      const double ex = (x * K.fx + K.cx) - (double)q.x;
      const double ey = (y * K.fy + K.cy) - (double)q.y;
      acc[27] += ex * ex + ey * ey;
      double jx[6], jy[6];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double dx0 = X * dR[j * 9 + 0] + Y * dR[j * 9 + 1] + Z * dR[j * 9 + 2];
        const double dy0 = X * dR[j * 9 + 3] + Y * dR[j * 9 + 4] + Z * dR[j * 9 + 5];
        const double dz0 = X * dR[j * 9 + 6] + Y * dR[j * 9 + 7] + Z * dR[j * 9 + 8];
        jx[j] = K.fx * (z * (dx0 - x * dz0));
        jy[j] = K.fy * (z * (dy0 - y * dz0));
      }
      jx[3] = K.fx * z; jx[4] = 0; jx[5] = K.fx * (-x * z);
      jy[3] = 0; jy[4] = K.fy * z; jy[5] = K.fy * (-y * z);
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) acc[k++] += jx[a] * jx[b] + jy[a] * jy[b];
#pragma unroll
      for (int a = 0; a < 6; a++) acc[21 + a] += jx[a] * ex + jy[a] * ey;
    };
#pragma unroll
    for (int k = 0; k < REF_RPT; k++)
      if (first + k * REF_CLUSTER * REF_THREADS < n) accumulate(Pk[k], qk[k]);
    for (int i = first + REF_RPT * REF_CLUSTER * REF_THREADS; i < n; i += REF_CLUSTER * REF_THREADS) {
      const int id = idx[i];
      accumulate(xyz[id], xy[id]);
    }
    // CTA reduction -> s_part[buf].  Warp level: a folding butterfly -- at step d a lane keeps the half of its
    // values its bit d selects and receives the partner's sums of that half, so 32 (padded) accumulators end as one
    // total per lane after 16+8+4+2+1 exchanges instead of 28 x 5.
    {
      double v16[16], v8[8], v4[4], v2[2], v1;
      {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const double lo = acc[k], hi = k + 16 < NACC ? acc[k + 16] : 0.;
          const double send = up ? lo : hi;
          v16[k] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool up = lane & 8;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const double send = up ? v16[k] : v16[k + 8];
          v8[k] = (up ? v16[k + 8] : v16[k]) + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool up = lane & 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const double send = up ? v8[k] : v8[k + 4];
          v4[k] = (up ? v8[k + 4] : v8[k]) + __shfl_xor_sync(0xffffffffu, send, 4);
        }
      }
      {
        const bool up = lane & 2;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const double send = up ? v4[k] : v4[k + 2];
          v2[k] = (up ? v4[k + 2] : v4[k]) + __shfl_xor_sync(0xffffffffu, send, 2);
        }
      }
      {
        const bool up = lane & 1;
        const double send = up ? v2[0] : v2[1];
        v1 = (up ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      // the lane's value is the warp total of accumulator (lane & 16) + (lane & 8) + ... = lane's bits, i.e. index `lane`
      if (lane < NACC) s_warp[w * NACC + lane] = v1;
    }
    __syncthreads();
    if (t < NACC) {
      double v = 0;
      for (int ww = 0; ww < REF_THREADS / 32; ww++) v += s_warp[ww * NACC + t];
      s_part[buf][t] = v;
    }
    cluster.sync();
    if (t < NACC) {
      double v = 0;
      for (unsigned r = 0; r < REF_CLUSTER; r++) v += *cluster.map_shared_rank(&s_part[buf][t], r);
      s_sum[t] = v;
    }
    __syncthreads();
    buf ^= 1;

    if (t == 0) {
      // LM step from prev_param with the current JtJ/JtErr and lambda
      auto lm_step = [&]() {
        const double lambda = c_lm_lambda[lambda_lg10 + 16];
        double A[36], x[6];
        for (int i = 0; i < 36; i++) A[i] = JtJ[i];
        for (int i = 0; i < 6; i++) A[i * 6 + i] *= 1. + lambda;
        solve6(A, JtErr, x);
        for (int i = 0; i < 6; i++) param[i] = prev_param[i] - x[i];
      };
      auto take_normal_equations = [&]() {   // CvLevMarq CALC_J: JtJ, JtErr at `param`, then step
        int k = 0;
        for (int a = 0; a < 6; a++)
          for (int b = a; b < 6; b++) {
            JtJ[a * 6 + b] = s_sum[k];
            JtJ[b * 6 + a] = s_sum[k];
            k++;
          }
        for (int a = 0; a < 6; a++) JtErr[a] = s_sum[21 + a];
        for (int i = 0; i < 6; i++) prev_param[i] = param[i];
        lm_step();
      };
      if (state == 0) {
        take_normal_equations();
        prev_err_norm = sqrt(s_sum[27]);   // iters == 0
        s_state = 1;
      } else {
        // CHECK_ERR at the candidate `param`
        err_norm = sqrt(s_sum[27]);
        bool retry = false;
        if (err_norm > prev_err_norm) {
          if (++lambda_lg10 <= 16) {
            lm_step();   // same JtJ / JtErr / prev_param, larger damping
            retry = true;
          }
        }
        if (!retry) {
          lambda_lg10 = lambda_lg10 - 1 > -16 ? lambda_lg10 - 1 : -16;
          double dn = 0, pn = 0;
          for (int i = 0; i < 6; i++) {
            dn += (param[i] - prev_param[i]) * (param[i] - prev_param[i]);
            pn += prev_param[i] * prev_param[i];
          }
          if (++iters >= max_iter || sqrt(dn) / sqrt(pn) < epsilon) {
            s_state = 2;
          } else {
            prev_err_norm = err_norm;
            take_normal_equations();   // the sums of this very pass are J, err at the accepted param
          }
        }
      }
      for (int i = 0; i < 6; i++) s_param[i] = param[i];
    }
    __syncthreads();
  }
  cluster.sync();   // nobody leaves while a peer may still read its shared memory
  if (t == 0 && rank == 0) {
    for (int i = 0; i < 6; i++) pose_out[i] = s_param[i];
    pose_out[6] = (double)iters;
    pose_out[7] = err_norm;
  }
}

int pnp_refine_launch(vo_ctx* c, const float3* xyz, const float2* xy, const int32_t* d_idx, const int* d_n_inl,
                      const double* d_models, const int* d_sel, double* d_pose) {
  Intrinsics K{c->p.fx, c->p.fy, c->p.cx, c->p.cy};
  {
    LaunchScope ls(c, VO_K_PNP_REFINE);
    pnp_refine_kernel<<<REF_CLUSTER, REF_THREADS, 0, c->stream>>>(xyz, xy, d_idx, d_n_inl, d_models, d_sel, K, d_pose);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
