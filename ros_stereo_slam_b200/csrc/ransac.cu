// ransac.cu -- K4/K6/K7: the two RANSAC estimators of the hot path, restructured for a GPU:
// every minimal-sample hypothesis is solved in parallel, every (hypothesis, point) pair is
// scored in parallel, and OpenCV's *sequential* acceptance rule (a hypothesis is accepted iff
// its inlier count beats every earlier one and lies before the adaptively shrinking iteration
// limit) is replayed afterwards by a prefix-max scan -- so the result is the one the serial
// loop of cv::RANSACPointSetRegistrator::run would return for the same sample list.
//
//   F-matrix : cv::findFundamentalMat(FM_RANSAC)  reference src/tracking.cpp:34,75
//   PnP      : cv::solvePnPRansac (EPnP-5)         reference src/keyFrameManagement.cpp:84,88
//
// All geometry is FP64 in OpenCV's operation order (cvmath.cuh / fmat7.cuh, --fmad=false);
// the residual is rounded to float and compared with (float)(thr*thr) like OpenCV does.
#include <stdlib.h>

#include "common.cuh"
#include "cvmath.cuh"
#include "fmat7.cuh"
#include "jacobi_warp.cuh"

namespace vo {

constexpr int PNP_STRIDE = 16;  // doubles per PnP model: rvec(3) tvec(3) R(9) pad
constexpr int F_STRIDE = 9;     // doubles per F model (3 per sample)

// ================================================================ hypothesis generation
// One WARP per minimal sample.  The solvers are long dependent FP64 chains (each Jacobi rotation
// carries 3 divides and 3 square roots), i.e. latency-bound: lane 0 runs the scalar parts, and
// the sweep phase of the big SVD (12x12 for EPnP, 7 rows x 9 for the 7-point solver) -- ~60 % of
// the chain -- is executed by the whole warp as a wavefront over independent row pairs
// (jacobi_warp.cuh), with bit-identical results.  The row matrix lives in shared memory.
constexpr int SOLVE_WARPS = 4;

__global__ void __launch_bounds__(SOLVE_WARPS * 32)
fmat_solve_kernel(const float2* __restrict__ m1, const float2* __restrict__ m2, const int32_t* __restrict__ samples,
                  int h, double* __restrict__ models, int32_t* __restrict__ counts) {
  __shared__ double s_v[SOLVE_WARPS][81];
  __shared__ double s_w[SOLVE_WARPS][8];
  __shared__ int s_ok[SOLVE_WARPS];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * SOLVE_WARPS + wid;
  if (s >= h) return;   // warp-uniform
  double* v = s_v[wid];
  double* W = s_w[wid];
  FmatNorm nm;
  if (lane == 0) {
    float a[14], b[14];
#pragma unroll
    for (int i = 0; i < 7; i++) {
      const int idx = samples[s * 7 + i];
      const float2 p = m1[idx], q = m2[idx];
      a[2 * i] = p.x; a[2 * i + 1] = p.y;
      b[2 * i] = q.x; b[2 * i + 1] = q.y;
    }
    const bool ok = fmat_7point_front(a, b, v, nm);
    s_ok[wid] = ok;
    if (ok)
      for (int i = 0; i < 7; i++) {
        double sd = 0;
        for (int k = 0; k < 9; k++) {
          const double t = v[i * 9 + k];
          sd += t * t;
        }
        W[i] = sd;
      }
  }
  __syncwarp();
  const bool ok = s_ok[wid] != 0;
  if (ok) jacobi_sweeps_warp<9, 7>(v, W, lane);
  if (lane == 0) {
    double F[27];
    int n = 0;
    if (ok) {
      double w[7];
      jacobi_svd<9, 7, 9, false, true>(v, w, v);   // norms, ordering, completion of rows 7 and 8
      n = fmat_7point_back(v, nm, F);
      if (n < 0 || n > 3) n = 0;
    }
    for (int k = 0; k < 3; k++) {
      counts[s * 3 + k] = k < n ? 0 : -1;
      if (k < n)
        for (int i = 0; i < 9; i++) models[(size_t)(s * 3 + k) * F_STRIDE + i] = F[k * 9 + i];
    }
  }
}

// EPnP-5 per warp.  Lane roles: lane 0 = scalar parts; all lanes = the 78 entries of MtM
// (each a sequential 10-term sum, as OpenCV's mulTransposed), the 12 row norms and the
// wavefront Jacobi sweeps; lanes 0..2 = the three independent beta initialisations N = 1,2,3
// with their Gauss-Newton polish and Horn alignment (one shared instruction stream).
struct PnpWarpSmem {
  double at[144];      // MtM -> U^T
  double W[12];
  double M[120];       // 10x12 design matrix; reused for l_6x10 (60) + rho (6) afterwards
  EpnpWork work;
  double var[3][13];   // per variant: reprojection error, R (9), t (3)
  double vAt[3][30], vV[3][25], vW[3][5];   // the three 6 x n systems of the beta initialisations (U^T rows, V^T, norms)
};

__global__ void __launch_bounds__(SOLVE_WARPS * 32)
pnp_solve_kernel(const float3* __restrict__ xyz, const float2* __restrict__ xy, const int32_t* __restrict__ samples,
                 int h, Intrinsics K, double* __restrict__ models, int32_t* __restrict__ counts) {
  __shared__ PnpWarpSmem sm_all[SOLVE_WARPS];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * SOLVE_WARPS + wid;
  if (s >= h) return;   // warp-uniform
  PnpWarpSmem& sm = sm_all[wid];
  double* at = sm.at;
  double* W = sm.W;
  if (lane == 0) {
    float obj[15], img[10];
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const int idx = samples[s * 5 + i];
      const float3 p = xyz[idx];
      const float2 q = xy[idx];
      obj[3 * i] = p.x; obj[3 * i + 1] = p.y; obj[3 * i + 2] = p.z;
      img[2 * i] = q.x; img[2 * i + 1] = q.y;
    }
    epnp5_front_M<false>(obj, img, K, sm.work, sm.M);
  }
  __syncwarp();
  // MtM = M^T M: upper triangle, one entry per lane and pass, mirrored (it is its own transpose,
  // so OpenCV's temp_a = MtM^T is `at` itself)
  for (int e = lane; e < 78; e += 32) {
    int i = 0, rem = e;
    while (rem >= 12 - i) {
      rem -= 12 - i;
      i++;
    }
    const int j = i + rem;
    double s0 = 0;
#pragma unroll
    for (int k = 0; k < 10; k++) s0 += sm.M[k * 12 + i] * sm.M[k * 12 + j];
    at[i * 12 + j] = s0;
    at[j * 12 + i] = s0;
  }
  __syncwarp();
  if (lane < 12) {
    double sd = 0;
#pragma unroll
    for (int k = 0; k < 12; k++) {
      const double t = at[lane * 12 + k];
      sd += t * t;
    }
    W[lane] = sd;
  }
  __syncwarp();
  jacobi_sweeps_warp<12, 12, true>(at, W, lane);
  double* l_6x10 = sm.M;
  double* rho = sm.M + 60;
  if (lane == 0) {
    double d[12];
    jacobi_svd<12, 12, 12, false, true>(at, d, at);   // norms, ordering, row normalisation -> U^T
    epnp_prepare(sm.work, at, l_6x10, rho);
  }
  __syncwarp();
  // the three beta initialisations: their 6 x n SVDs run as wavefronts side by side (lane group g = variant g + 1), the
  // rest of a variant on one lane each (three lanes, one instruction stream)
  if (lane < 3) {
    epnp_variant_system(lane + 1, l_6x10, sm.vAt[lane]);
    jacobi_rt_init<6, 5>(sm.vAt[lane], sm.vW[lane], sm.vV[lane], epnp_variant_ncol(lane + 1));
  }
  __syncwarp();
  jacobi_sweeps_groups<6, 5>(&sm.vAt[0][0], 30, &sm.vW[0][0], 5, &sm.vV[0][0], 25, lane < 24 ? epnp_variant_ncol((lane >> 3) + 1) : 0, lane);
  __syncwarp();
  if (lane < 3) {
    const int ncol = epnp_variant_ncol(lane + 1);
    double sw[5], bx[5] = {0, 0, 0, 0, 0};
    jacobi_rt_finish<6, 5>(sm.vAt[lane], sw, sm.vV[lane], ncol);
    svd_backsubst_6xn(sm.vAt[lane], sw, sm.vV[lane], rho, bx, ncol);
    double betas[4], R[3][3], t[3];
    const double rep = epnp_variant_back(lane + 1, sm.work, K, at, l_6x10, rho, bx, betas, R, t);
    double* o = sm.var[lane];
    o[0] = rep;
    for (int i = 0; i < 3; i++) {
      o[10 + i] = t[i];
      for (int j = 0; j < 3; j++) o[1 + i * 3 + j] = R[i][j];
    }
  }
  __syncwarp();
  if (lane == 0) {
    int N = 1;
    if (sm.var[1][0] < sm.var[0][0]) N = 2;
    if (sm.var[2][0] < sm.var[N - 1][0]) N = 3;
    const double* o = sm.var[N - 1];
    double R[9], t[3], rvec[3], R2[9];
    for (int i = 0; i < 9; i++) R[i] = o[1 + i];
    for (int i = 0; i < 3; i++) t[i] = o[10 + i];
    rodrigues_mat2vec(R, rvec);       // the RANSAC model is (rvec, tvec) ...
    rodrigues_vec2mat(rvec, R2);      // ... and projectPoints converts it back
    double* m = models + (size_t)s * PNP_STRIDE;
    for (int i = 0; i < 3; i++) {
      m[i] = rvec[i];
      m[3 + i] = t[i];
    }
    for (int i = 0; i < 9; i++) m[6 + i] = R2[i];
    m[15] = 0;
    counts[s] = 0;
  }
}

// debug: one EPnP on the device with intermediates (parity investigations only)
__global__ void epnp_debug_kernel(const float* obj, const float* img, Intrinsics K, double* dbg) {
  double R[9], t[3];
  epnp5<false>(obj, img, K, R, t, dbg);
  for (int i = 0; i < 9; i++) dbg[420 + i] = R[i];
  for (int i = 0; i < 3; i++) dbg[429 + i] = t[i];
}

int epnp_debug_launch(vo_ctx* c, const float* d_obj, const float* d_img, double* d_dbg) {
  Intrinsics K{c->p.fx, c->p.fy, c->p.cx, c->p.cy};
  epnp_debug_kernel<<<1, 1, 0, c->stream>>>(d_obj, d_img, K, d_dbg);
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ================================================================ sequential acceptance replay
// cv::RANSACUpdateNumIters
__device__ int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.);
  p = fmin(p, 1.);
  ep = fmax(ep, 0.);
  ep = fmin(ep, 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

constexpr int SEL_THREADS = 1024;

// What the selection needs besides the counts; passed by value to the score kernels when the selection is fused
// into their last CTA (sel == nullptr: no fused selection).
struct SelectArgs {
  int n_samples, mps, model_points, n_points;
  double conf;
  int max_iters;
  int* sel;
  unsigned* done;   // CTA completion counter of the score launch (zero between launches)
};

// counts: flattened [sample][model] inlier counts (-1 = no such model).  One CTA of NT threads.
// A model is a *candidate* iff its count exceeds every earlier count and modelPoints-1
// (prefix max); candidates are then replayed in order with the adaptive iteration limit.
// The counts are read through L2 (__ldcg): the fused caller's CTA may hold stale lines of them in L1.
template <int NT>
__device__ __forceinline__ void select_block(const int32_t* __restrict__ counts, int n_samples, int mps, int model_points,
                                             int n_points, double conf, int max_iters, int* __restrict__ sel) {
  __shared__ int s_excl[NT];
  __shared__ int s_warp[32];
  __shared__ unsigned s_cand[32];
  const int t = threadIdx.x;
  const int total = n_samples * mps;
  const int chunk = (total + NT - 1) / NT;
  const int beg = t * chunk, end = min(beg + chunk, total);
  int mx = -1;
  for (int i = beg; i < end; i++) mx = max(mx, __ldcg(counts + i));
  // inclusive prefix max over threads
  int incl = mx;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((t & 31) >= d) incl = max(incl, v);
  }
  if ((t & 31) == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    int w = t < NT / 32 ? s_warp[t] : -1;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, d);
      if (t >= d) wi = max(wi, v);
    }
    // exclusive over warps
    int ex = __shfl_up_sync(0xffffffffu, wi, 1);
    s_warp[t] = t == 0 ? -1 : ex;
  }
  __syncthreads();
  int ex_lane = __shfl_up_sync(0xffffffffu, incl, 1);
  if ((t & 31) == 0) ex_lane = -1;
  const int floor0 = model_points - 1;
  const int excl = max(max(s_warp[t >> 5], ex_lane), floor0);
  s_excl[t] = excl;
  const unsigned cand = __ballot_sync(0xffffffffu, mx > excl);
  if ((t & 31) == 0) s_cand[t >> 5] = cand;
  __syncthreads();
  if (t == 0) {
    int niters = max(max_iters, 1);
    int best = -1, best_count = 0, n_rec = 0;
    int open_sample = -1;  // OpenCV tests `iter < niters` once per sample: all models of an
                           // admitted sample are scored even if the first one shrinks niters
    bool stop = false;
    for (int w = 0; w < NT / 32 && !stop; w++) {
      unsigned bits = s_cand[w];
      while (bits && !stop) {
        const int l = __ffs(bits) - 1;
        bits &= bits - 1;
        const int th = w * 32 + l;
        int run = s_excl[th];
        const int b2 = th * chunk, e2 = min(b2 + chunk, total);
        for (int i = b2; i < e2; i++) {
          const int g = __ldcg(counts + i);
          if (g > run) {
            const int smp = i / mps;
            if (smp != open_sample) {
              if (smp >= niters) {
                stop = true;
                break;
              }
              open_sample = smp;
            }
            run = g;
            best = i;
            best_count = g;
            n_rec++;
            niters = ransac_update_num_iters(conf, (double)(n_points - g) / n_points, model_points, niters);
          }
        }
      }
    }
    sel[0] = best;
    sel[1] = niters;
    sel[2] = best_count;
    sel[3] = n_rec;
  }
}

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(const int32_t* __restrict__ counts, int n_samples, int mps, int model_points, int n_points, double conf,
              int max_iters, int* __restrict__ sel, const int* __restrict__ n_dev) {
  if (n_dev) n_points = min(n_points, *n_dev);
  select_block<SEL_THREADS>(counts, n_samples, mps, model_points, n_points, conf, max_iters, sel);
}

// Tail of a score kernel with a fused selection: the CTA that finishes last (device counter) replays the
// acceptance sequence over the now complete counts.  Every thread fences its own count updates first.
template <int NT>
__device__ __forceinline__ void score_tail_select(const int32_t* __restrict__ counts, const SelectArgs& sa, int n_points) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    s_last = atomicAdd(sa.done, 1u) == total - 1;
    if (s_last) *sa.done = 0;   // ready for the next launch (stream order)
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  select_block<NT>(counts, sa.n_samples, sa.mps, sa.model_points, n_points, sa.conf, sa.max_iters, sa.sel);
}

// ================================================================ scoring
__device__ __forceinline__ float fmat_err(const double* F, float2 p1, float2 p2) {
  const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
  double a = F[0] * x1 + F[1] * y1 + F[2];
  double b = F[3] * x1 + F[4] * y1 + F[5];
  double c = F[6] * x1 + F[7] * y1 + F[8];
  const double s2 = 1. / (a * a + b * b);
  const double d2 = x2 * a + y2 * b + c;
  a = F[0] * x2 + F[3] * y2 + F[6];
  b = F[1] * x2 + F[4] * y2 + F[7];
  c = F[2] * x2 + F[5] * y2 + F[8];
  const double s1 = 1. / (a * a + b * b);
  const double d1 = x1 * a + y1 * b + c;
  const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
  return (float)((e1 < e2) ? e2 : e1);  // std::max(e1, e2)
}

__device__ __forceinline__ float pnp_err(const double* m /*PNP_STRIDE*/, const Intrinsics& K, float3 P, float2 q) {
  const double* R = m + 6;
  const double* t = m + 3;
  const double X = P.x, Y = P.y, Z = P.z;
  double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
  double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
  double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
  z = z ? 1. / z : 1;
  x *= z;
  y *= z;
  const float u = (float)(x * K.fx + K.cx);
  const float v = (float)(y * K.fy + K.cy);
  const float dx = __fsub_rn(q.x, u), dy = __fsub_rn(q.y, v);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

constexpr int SCORE_TPB = 256;
constexpr int SCORE_HB = 16;  // hypotheses per CTA (staged in shared memory)

// grid.x = point tiles, grid.y = hypothesis batches.  Each thread keeps its correspondence in
// registers and walks the batch; per-hypothesis counts are reduced warp (ballot) -> CTA
// (shared atomics) -> global (one atomic per CTA and hypothesis).
__global__ void __launch_bounds__(SCORE_TPB)
fmat_score_kernel(const float2* __restrict__ m1, const float2* __restrict__ m2, int n, const double* __restrict__ models,
                  int32_t* __restrict__ counts, int n_models, float thr2, const int* __restrict__ limit,
                  const int* __restrict__ n_dev, SelectArgs sa) {
  if (n_dev) n = min(n, *n_dev);
  __shared__ double sF[SCORE_HB * F_STRIDE];
  __shared__ int sCnt[SCORE_HB];
  __shared__ int sValid[SCORE_HB];
  const int m0 = blockIdx.y * SCORE_HB;
  if (limit && m0 >= *limit * 3) return;   // (never together with a fused selection)
  for (int i = threadIdx.x; i < SCORE_HB * F_STRIDE; i += blockDim.x) {
    const int m = m0 + i / F_STRIDE;
    sF[i] = m < n_models ? models[(size_t)m0 * F_STRIDE + i] : 0.;
  }
  if (threadIdx.x < SCORE_HB) {
    const int m = m0 + threadIdx.x;
    sCnt[threadIdx.x] = 0;
    sValid[threadIdx.x] = (m < n_models) && counts[m] >= 0;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < n;
  float2 p1 = make_float2(0, 0), p2 = make_float2(0, 0);
  if (act) {
    p1 = m1[i];
    p2 = m2[i];
  }
#pragma unroll 1
  for (int k = 0; k < SCORE_HB; k++) {
    if (!sValid[k]) continue;
    const bool in = act && (fmat_err(sF + k * F_STRIDE, p1, p2) <= thr2);
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&sCnt[k], __popc(b));
  }
  __syncthreads();
  if (threadIdx.x < SCORE_HB && sValid[threadIdx.x] && sCnt[threadIdx.x])
    atomicAdd(&counts[m0 + threadIdx.x], sCnt[threadIdx.x]);
  if (sa.sel) score_tail_select<SCORE_TPB>(counts, sa, n);
}

__global__ void __launch_bounds__(SCORE_TPB)
pnp_score_kernel(const float3* __restrict__ xyz, const float2* __restrict__ xy, int n, const double* __restrict__ models,
                 int32_t* __restrict__ counts, int n_models, Intrinsics K, float thr2, const int* __restrict__ limit,
                 const int* __restrict__ n_dev, SelectArgs sa) {
  if (n_dev) n = min(n, *n_dev);
  __shared__ double sM[SCORE_HB * PNP_STRIDE];
  __shared__ int sCnt[SCORE_HB];
  const int m0 = blockIdx.y * SCORE_HB;
  if (limit && m0 >= *limit) return;
  for (int i = threadIdx.x; i < SCORE_HB * PNP_STRIDE; i += blockDim.x) {
    const int m = m0 + i / PNP_STRIDE;
    sM[i] = m < n_models ? models[(size_t)m0 * PNP_STRIDE + i] : 0.;
  }
  if (threadIdx.x < SCORE_HB) sCnt[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < n;
  float3 P = make_float3(0, 0, 1);
  float2 q = make_float2(0, 0);
  if (act) {
    P = xyz[i];
    q = xy[i];
  }
  const int kmax = min(SCORE_HB, n_models - m0);
#pragma unroll 1
  for (int k = 0; k < kmax; k++) {
    const bool in = act && (pnp_err(sM + k * PNP_STRIDE, K, P, q) <= thr2);
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&sCnt[k], __popc(b));
  }
  __syncthreads();
  if (threadIdx.x < kmax && sCnt[threadIdx.x]) atomicAdd(&counts[m0 + threadIdx.x], sCnt[threadIdx.x]);
  if (sa.sel) score_tail_select<SCORE_TPB>(counts, sa, n);
}

// ================================================================ best-model inlier mask
__global__ void fmat_mask_kernel(const float2* __restrict__ m1, const float2* __restrict__ m2, int n,
                                 const double* __restrict__ models, const int* __restrict__ sel, float thr2,
                                 uint8_t* __restrict__ mask, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int best = sel[0];
  uint8_t v = 0;
  if (best >= 0) v = fmat_err(models + (size_t)best * F_STRIDE, m1[i], m2[i]) <= thr2;
  mask[i] = v;
}

__global__ void pnp_mask_kernel(const float3* __restrict__ xyz, const float2* __restrict__ xy, int n,
                                const double* __restrict__ models, const int* __restrict__ sel, Intrinsics K,
                                float thr2, uint8_t* __restrict__ mask, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int best = sel[0];
  uint8_t v = 0;
  if (best >= 0) v = pnp_err(models + (size_t)best * PNP_STRIDE, K, xyz[i], xy[i]) <= thr2;
  mask[i] = v;
}

// ================================================================ best-model mask + ordered compaction, one launch
// The fused chains' "mask -> compact" pair: every thread evaluates the best model on its CP_ITEMS consecutive
// correspondences, the tile is compacted with the shared look-back core (common.cuh).  PNP = false: F-matrix error on
// (a_in, b_in), survivors of a_in / b_in / c_in (optional) are copied; PNP = true: reprojection error of (c_in, a_in),
// the survivors' indices are written.
template <bool PNP>
__global__ void __launch_bounds__(CP_THREADS)
mask_compact_kernel(const float2* __restrict__ a_in, const float2* __restrict__ b_in, const float3* __restrict__ c_in, int n,
                    const double* __restrict__ models, const int* __restrict__ sel, Intrinsics K, float thr2,
                    uint8_t* __restrict__ mask, float2* __restrict__ a_out, float2* __restrict__ b_out,
                    float3* __restrict__ c_out, int32_t* __restrict__ idx_out, int* __restrict__ count_out,
                    volatile unsigned long long* tile_state, unsigned* epoch_ctr, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  constexpr int STRIDE = PNP ? PNP_STRIDE : F_STRIDE;
  __shared__ double sM[STRIDE];
  const int best = sel[0];
  const int t = threadIdx.x;
  if (best >= 0 && t < STRIDE) sM[t] = models[(size_t)best * STRIDE + t];
  __syncthreads();
  const int beg = blockIdx.x * CP_TILE + t * CP_ITEMS;
  unsigned long long bits = 0;
  int cnt = 0;
  if (best >= 0) {
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) {
      const int i = beg + k;
      if (i < n) {
        bool in;
        if (PNP) in = pnp_err(sM, K, c_in[i], a_in[i]) <= thr2;
        else in = fmat_err(sM, a_in[i], b_in[i]) <= thr2;
        bits |= (unsigned long long)in << (8 * k);
        cnt += in;
      }
    }
  }
  if (beg + CP_ITEMS <= n) {
    *reinterpret_cast<unsigned long long*>(mask + beg) = bits;
  } else {
    for (int k = 0; k < CP_ITEMS; k++)
      if (beg + k < n) mask[beg + k] = (uint8_t)(bits >> (8 * k));
  }
  int pos = compact_tile_offset(cnt, count_out, tile_state, epoch_ctr);
#pragma unroll
  for (int k = 0; k < CP_ITEMS; k++) {
    if ((bits >> (8 * k)) & 1) {
      const int i = beg + k;
      if (PNP) {
        idx_out[pos] = i;
      } else {
        a_out[pos] = a_in[i];
        b_out[pos] = b_in[i];
        if (c_in) c_out[pos] = c_in[i];
      }
      pos++;
    }
  }
}

// ================================================================ device-side sampling
// RANSACPointSetRegistrator::getSubset for all H samples of a chunk at once.  OpenCV consumes ONE
// random stream sequentially: sample s starts where sample s-1 stopped, a duplicate index costs one
// extra draw, a subset rejected by the collinearity check (F-matrix only) costs a whole new subset.
// The stream itself is fixed (seed 2^64-1) and precomputed in `raw`.  Every thread owns one sample,
// speculates that all earlier samples consumed exactly M values, computes its own consumption from
// that start, and a block prefix sum yields the true start positions; threads whose start moved
// recompute.  Each round fixes at least the first wrong start, and exceptions are rare (a few per
// thousand samples; a few per 48 on integer grids), so this converges in a handful of rounds.
// Latency: the random stream window the block will consume (M*H values + slack) is staged into
// shared memory with coalesced loads, so the draws of a subset are shared-memory reads instead of a
// chain of dependent global loads, and the 2 x M points of the collinearity check are fetched with
// independent loads before the 21 pair tests run out of registers.
constexpr int SAMPLE_SLACK = 256;

template <int M, bool CHECK>
__global__ void __launch_bounds__(1024)
sample_kernel(const uint32_t* __restrict__ raw, int raw_len, const int* __restrict__ n_dev, int n_max,
              const float2* __restrict__ m1, const float2* __restrict__ m2, int H, int32_t* __restrict__ out,
              int* __restrict__ flag) {
  extern __shared__ uint32_t s_raw[];   // M*H + SAMPLE_SLACK values of the stream
  __shared__ int s_warp[32];
  __shared__ int s_bad;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int n = n_max;
  if (n_dev) n = min(n, *n_dev);
  const int staged = min(M * H + SAMPLE_SLACK, raw_len);
  for (int i = t; i < staged; i += blockDim.x) s_raw[i] = raw[i];
  if (t == 0) s_bad = 0;
  __syncthreads();
  if (n < M + 1) {   // too few points for this estimator's RANSAC path (host decides what to do)
    // the solve kernels of a fused chain still run on `out` before the host sees the flag: give them indices they
    // may dereference (index 0 of buffers that always hold >= 32 elements), never stale memory
    for (int i = t; i < M * H; i += blockDim.x) out[i] = 0;
    if (t == 0) *flag = 2;
    return;
  }
  const bool active = t < H;
  int pos = M * t, used = M, idx[M];
  bool dirty = true;
  for (int round = 0; round <= H + 1; round++) {
    if (active && dirty) {
      int p = pos;
      bool ok = false;
      for (int attempt = 0; attempt < 10000 && !ok; attempt++) {
#pragma unroll
        for (int i = 0; i < M; i++) {
          int v;
          for (;;) {
            if (p >= raw_len) { s_bad = 1; v = 0; break; }
            const uint32_t rv = p < staged ? s_raw[p] : raw[p];
            p++;
            v = (int)(rv % (unsigned)n);
            bool dup = false;
#pragma unroll
            for (int k = 0; k < M; k++) dup |= (k < i && idx[k] == v);
            if (!dup) break;
          }
          idx[i] = v;
        }
        ok = true;
        if (CHECK && !s_bad) {
          // haveCollinearPoints(ms1) || haveCollinearPoints(ms2): last point vs every earlier pair
#pragma unroll
          for (int set = 0; set < 2; set++) {
            const float2* pts = set ? m2 : m1;
            float2 q[M];
#pragma unroll
            for (int i = 0; i < M; i++) q[i] = pts[idx[i]];
            const float2 pi = q[M - 1];
            bool col = false;
#pragma unroll
            for (int j = 0; j < M - 1; j++) {
              const double dx1 = (double)q[j].x - (double)pi.x, dy1 = (double)q[j].y - (double)pi.y;
#pragma unroll
              for (int k = 0; k < M - 1; k++) {
                if (k >= j) continue;
                const double dx2 = (double)q[k].x - (double)pi.x, dy2 = (double)q[k].y - (double)pi.y;
                col |= fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2));
              }
            }
            if (col) ok = false;
          }
        }
        if (s_bad) break;
      }
      if (!ok) s_bad = 1;
      used = p - pos;
    }
    // exclusive prefix sum of `used` over the block -> true start positions
    int incl = active ? used : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
      int x = s_warp[lane], xi = x;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, xi, d);
        if (lane >= d) xi += v;
      }
      s_warp[lane] = xi - x;
    }
    __syncthreads();
    const int newpos = s_warp[w] + incl - (active ? used : 0);
    dirty = active && newpos != pos;
    pos = newpos;
    const int any = __syncthreads_or(dirty ? 1 : 0);
    if (!any || s_bad) break;
  }
  if (active)
    for (int i = 0; i < M; i++) out[t * M + i] = idx[i];
  if (t == 0) *flag = s_bad ? 1 : 0;
}

int sample_launch(vo_ctx* c, int model_points, const float2* m1, const float2* m2, int n_max, int h, int32_t* d_samples,
                  int* d_flag) {
  if (h <= 0 || h > 1024) return VO_ERR_INVALID_ARG;
  {
    LaunchScope ls(c, VO_K_SELECT);
    // threads: one per sample, rounded up to whole warps (the block-wide scans assume full warps)
    const int threads = std::min(1024, std::max(64, div_up(h, 32) * 32));
    const size_t smem = (size_t)(model_points * h + SAMPLE_SLACK) * sizeof(uint32_t);
    if (model_points == 7)
      sample_kernel<7, true><<<1, threads, smem, c->stream>>>(c->d_rng, RNG_LEN, c->n_dev, n_max, m1, m2, h, d_samples, d_flag);
    else
      sample_kernel<5, false><<<1, threads, smem, c->stream>>>(c->d_rng, RNG_LEN, c->n_dev, n_max, nullptr, nullptr, h,
                                                             d_samples, d_flag);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ================================================================ launchers
static Intrinsics intr(const vo_ctx* c) { return Intrinsics{c->p.fx, c->p.fy, c->p.cx, c->p.cy}; }

int fmat_solve_launch(vo_ctx* c, const float2* m1, const float2* m2, const int32_t* d_samples, int h, double* d_models,
                      int32_t* d_counts) {
  if (h <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_FMAT_SOLVE);
    fmat_solve_kernel<<<div_up(h, SOLVE_WARPS), SOLVE_WARPS * 32, 0, c->stream>>>(m1, m2, d_samples, h, d_models, d_counts);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int fmat_score_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, int32_t* d_counts,
                      int h, float thr2) {
  if (h <= 0 || n <= 0) return VO_OK;
  dim3 g(div_up(n, SCORE_TPB), div_up(h * 3, SCORE_HB));
  {
    LaunchScope ls(c, VO_K_FMAT_SCORE);
    fmat_score_kernel<<<g, SCORE_TPB, 0, c->stream>>>(m1, m2, n, d_models, d_counts, h * 3, thr2, nullptr, c->n_dev, SelectArgs{});
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// scoring with the selection (select_launch's arguments) fused into the last CTA
int fmat_score_select_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, int32_t* d_counts,
                             int h, float thr2, double conf, int max_iters, int* d_sel) {
  if (h <= 0 || n <= 0) return select_launch(c, d_counts, h, 3, 7, n, conf, max_iters, d_sel);
  dim3 g(div_up(n, SCORE_TPB), div_up(h * 3, SCORE_HB));
  {
    LaunchScope ls(c, VO_K_FMAT_SCORE);
    fmat_score_kernel<<<g, SCORE_TPB, 0, c->stream>>>(m1, m2, n, d_models, d_counts, h * 3, thr2, nullptr, c->n_dev,
                                                     SelectArgs{h, 3, 7, n, conf, max_iters, d_sel, c->d_epoch + 1});
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int fmat_mask_compact_launch(vo_ctx* c, const float2* m1, const float2* m2, const float3* xyz, int n, const double* d_models,
                             const int* d_sel, float thr2, uint8_t* d_mask, float2* o1, float2* o2, float3* oxyz,
                             int count_slot) {
  if (n <= 0) {
    VO_CUDA(cudaMemsetAsync(c->d_count + count_slot, 0, sizeof(int), c->stream));
    return VO_OK;
  }
  {
    LaunchScope ls(c, VO_K_COMPACT);
    mask_compact_kernel<false><<<div_up(n, CP_TILE), CP_THREADS, 0, c->stream>>>(
        m1, m2, xyz, n, d_models, d_sel, intr(c), thr2, d_mask, o1, o2, oxyz, nullptr, c->d_count + count_slot,
        c->d_tile_state, c->d_epoch, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int fmat_mask_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, const int* d_sel,
                     float thr2, uint8_t* d_mask) {
  if (n <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_FMAT_SCORE);
    fmat_mask_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(m1, m2, n, d_models, d_sel, thr2, d_mask, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int pnp_solve_launch(vo_ctx* c, const float3* xyz, const float2* xy, const int32_t* d_samples, int h, double* d_models,
                     int32_t* d_counts) {
  if (h <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_PNP_SOLVE);
    pnp_solve_kernel<<<div_up(h, SOLVE_WARPS), SOLVE_WARPS * 32, 0, c->stream>>>(xyz, xy, d_samples, h, intr(c), d_models,
                                                                       d_counts);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int pnp_score_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, int32_t* d_counts,
                     int h, float thr2) {
  if (h <= 0 || n <= 0) return VO_OK;
  dim3 g(div_up(n, SCORE_TPB), div_up(h, SCORE_HB));
  {
    LaunchScope ls(c, VO_K_PNP_SCORE);
    pnp_score_kernel<<<g, SCORE_TPB, 0, c->stream>>>(xyz, xy, n, d_models, d_counts, h, intr(c), thr2, nullptr, c->n_dev,
                                                    SelectArgs{});
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int pnp_score_select_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, int32_t* d_counts,
                            int h, float thr2, double conf, int max_iters, int* d_sel) {
  if (h <= 0 || n <= 0) return select_launch(c, d_counts, h, 1, 5, n, conf, max_iters, d_sel);
  dim3 g(div_up(n, SCORE_TPB), div_up(h, SCORE_HB));
  {
    LaunchScope ls(c, VO_K_PNP_SCORE);
    pnp_score_kernel<<<g, SCORE_TPB, 0, c->stream>>>(xyz, xy, n, d_models, d_counts, h, intr(c), thr2, nullptr, c->n_dev,
                                                    SelectArgs{h, 1, 5, n, conf, max_iters, d_sel, c->d_epoch + 1});
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int pnp_mask_compact_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, const int* d_sel,
                            float thr2, uint8_t* d_mask, int32_t* d_idx, int count_slot) {
  if (n <= 0) {
    VO_CUDA(cudaMemsetAsync(c->d_count + count_slot, 0, sizeof(int), c->stream));
    return VO_OK;
  }
  {
    LaunchScope ls(c, VO_K_COMPACT);
    mask_compact_kernel<true><<<div_up(n, CP_TILE), CP_THREADS, 0, c->stream>>>(
        xy, nullptr, xyz, n, d_models, d_sel, intr(c), thr2, d_mask, nullptr, nullptr, nullptr, d_idx, c->d_count + count_slot,
        c->d_tile_state, c->d_epoch, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int pnp_mask_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, const int* d_sel,
                    float thr2, uint8_t* d_mask) {
  if (n <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_PNP_SCORE);
    pnp_mask_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(xyz, xy, n, d_models, d_sel, intr(c), thr2, d_mask, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ================================================================ four and five points
// cv::solvePnPRansac does not run RANSAC when the minimal sample is the whole set (calib3d solvepnp.cpp,
// `model_points == npoints`): five points -> solvePnP(EPNP) on all of them, four points -> solvePnP(P3P) (three points
// give up to four poses, the fourth picks the one with the smallest reprojection error); the inlier list is every
// point and no refinement follows.  The five-point case is pnp_solve_kernel on the sample (0..4).  For four points:
// Grunert's formulation with the distance ratios u = s2/s1, v = s3/s1, which leaves one quartic in v (coefficients by
// polynomial arithmetic, roots by Durand-Kerner + Newton); the pose follows from the two orthonormal frames spanned
// by the triangle in world and camera coordinates.  OpenCV's own P3P code is a different formulation of the same
// equations: the poses agree to rounding, not bit for bit (tests: 1e-6 rad / 1e-6 m against cv2).
__device__ void poly_mul(const double* a, int na, const double* b, int nb, double* out) {   // ascending powers
  for (int i = 0; i <= na + nb; i++) out[i] = 0;
  for (int i = 0; i <= na; i++)
    for (int j = 0; j <= nb; j++) out[i + j] += a[i] * b[j];
}

__device__ int quartic_real_roots(const double c[5], double roots[4]) {
  // degree handling: a vanishing leading coefficient (degenerate geometry) lowers the degree
  int deg = 4;
  double scale = 0;
  for (int i = 0; i <= 4; i++) scale = fmax(scale, fabs(c[i]));
  if (scale == 0) return 0;
  while (deg > 0 && fabs(c[deg]) < 1e-14 * scale) deg--;
  if (deg == 0) return 0;
  double a[5];
  for (int i = 0; i <= deg; i++) a[i] = c[i] / c[deg];      // monic
  double zr[4], zi[4];
  // Durand-Kerner from points on a circle of the Cauchy bound
  double bound = 0;
  for (int i = 0; i < deg; i++) bound = fmax(bound, fabs(a[i]));
  bound += 1;
  for (int k = 0; k < deg; k++) {
    const double ang = 0.4 + 6.283185307179586 * k / deg;
    zr[k] = 0.5 * bound * cos(ang);
    zi[k] = 0.5 * bound * sin(ang);
  }
  for (int it = 0; it < 200; it++) {
    double change = 0;
    for (int k = 0; k < deg; k++) {
      double pr = 1, pi = 0;                      // p(z_k), Horner, monic
      for (int i = deg - 1; i >= 0; i--) {
        const double tr = pr * zr[k] - pi * zi[k] + a[i], ti = pr * zi[k] + pi * zr[k];
        pr = tr;
        pi = ti;
      }
      double qr = 1, qi = 0;                      // prod (z_k - z_j)
      for (int j = 0; j < deg; j++)
        if (j != k) {
          const double dr = zr[k] - zr[j], di = zi[k] - zi[j];
          const double tr = qr * dr - qi * di, ti = qr * di + qi * dr;
          qr = tr;
          qi = ti;
        }
      const double den = qr * qr + qi * qi;
      if (den == 0) continue;
      const double wr = (pr * qr + pi * qi) / den, wi = (pi * qr - pr * qi) / den;
      zr[k] -= wr;
      zi[k] -= wi;
      change = fmax(change, fabs(wr) + fabs(wi));
    }
    if (change < 1e-15 * bound) break;
  }
  int n = 0;
  for (int k = 0; k < deg; k++) {
    if (fabs(zi[k]) > 1e-7 * (1 + fabs(zr[k]))) continue;
    double x = zr[k];
    for (int it = 0; it < 8; it++) {               // Newton polish on the real polynomial
      double f = 1, d = 0;
      for (int i = deg - 1; i >= 0; i--) {
        d = d * x + f;
        f = f * x + a[i];
      }
      if (d == 0) break;
      x -= f / d;
    }
    roots[n++] = x;
  }
  return n;
}

// up to 4 poses (R row-major, t) from three correspondences; f = unit bearing vectors, P = world points
__device__ int p3p_solve3(const double P[3][3], const double f[3][3], double Rs[4][9], double ts[4][3]) {
  auto sub = [](const double* a, const double* b, double* o) { for (int i = 0; i < 3; i++) o[i] = a[i] - b[i]; };
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  auto cross = [](const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  double d[3];
  sub(P[1], P[2], d); const double a2 = dot(d, d);
  sub(P[0], P[2], d); const double b2 = dot(d, d);
  sub(P[0], P[1], d); const double c2 = dot(d, d);
  if (a2 == 0 || b2 == 0 || c2 == 0) return 0;
  const double ca = dot(f[1], f[2]), cb = dot(f[0], f[2]), cg = dot(f[0], f[1]);
  const double K1 = (a2 - c2) / b2, K2 = c2 / b2;
  // u = N(v) / D(v);  D^2 + N^2 - 2 N D cos(gamma) - K2 (1 + v^2 - 2 v cos(beta)) D^2 = 0
  const double N[3] = {1 + K1, -2 * K1 * cb, K1 - 1};
  const double D[2] = {2 * cg, -2 * ca};
  const double Q[3] = {1, -2 * cb, 1};
  double D2[3], N2[5], ND[4], QD2[5], c[5];
  poly_mul(D, 1, D, 1, D2);
  poly_mul(N, 2, N, 2, N2);
  poly_mul(N, 2, D, 1, ND);
  poly_mul(Q, 2, D2, 2, QD2);
  for (int i = 0; i < 5; i++) c[i] = N2[i] - K2 * QD2[i] + (i < 3 ? D2[i] : 0) - (i < 4 ? 2 * cg * ND[i] : 0);
  double v[4];
  const int nr = quartic_real_roots(c, v);
  int ns = 0;
  for (int r = 0; r < nr && ns < 4; r++) {
    const double vv = v[r];
    const double den = D[0] + D[1] * vv;
    const double q = 1 + vv * vv - 2 * vv * cb;
    if (fabs(den) < 1e-12 || q <= 0) continue;
    const double uu = (N[0] + N[1] * vv + N[2] * vv * vv) / den;
    const double s1 = sqrt(b2 / q);
    const double s2 = uu * s1, s3 = vv * s1;
    if (!(s2 > 0) || !(s3 > 0)) continue;
    bool dup = false;                              // double roots of the quartic give the same pose twice
    for (int k = 0; k < r; k++) dup |= fabs(v[k] - vv) < 1e-9 * (1 + fabs(vv));
    if (dup) continue;
    double C[3][3];
    for (int i = 0; i < 3; i++) {
      C[0][i] = s1 * f[0][i];
      C[1][i] = s2 * f[1][i];
      C[2][i] = s3 * f[2][i];
    }
    // orthonormal frames of the two triangles
    double ep[3][3], ec[3][3], t1[3], t2[3];
    for (int w = 0; w < 2; w++) {
      const double(*X)[3] = w == 0 ? P : C;
      double(*e)[3] = w == 0 ? ep : ec;
      sub(X[1], X[0], t1);
      sub(X[2], X[0], t2);
      double n1 = sqrt(dot(t1, t1));
      for (int i = 0; i < 3; i++) e[0][i] = t1[i] / n1;
      cross(e[0], t2, e[2]);
      double n3 = sqrt(dot(e[2], e[2]));
      if (n3 == 0) return ns;                      // collinear points
      for (int i = 0; i < 3; i++) e[2][i] /= n3;
      cross(e[2], e[0], e[1]);
    }
    // R = Ec^T Ep  (maps world frame vectors to camera frame vectors)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rs[ns][i * 3 + j] = ec[0][i] * ep[0][j] + ec[1][i] * ep[1][j] + ec[2][i] * ep[2][j];
    for (int i = 0; i < 3; i++)
      ts[ns][i] = C[0][i] - (Rs[ns][i * 3] * P[0][0] + Rs[ns][i * 3 + 1] * P[0][1] + Rs[ns][i * 3 + 2] * P[0][2]);
    ns++;
  }
  return ns;
}

// model (PNP_STRIDE doubles, layout of pnp_solve_kernel) of the first four points; counts[0] = 0 or -1
__global__ void pnp_p3p4_kernel(const float3* __restrict__ xyz, const float2* __restrict__ xy, Intrinsics K,
                                double* __restrict__ models, int32_t* __restrict__ counts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double P[3][3], f[3][3];
  for (int i = 0; i < 3; i++) {
    P[i][0] = xyz[i].x; P[i][1] = xyz[i].y; P[i][2] = xyz[i].z;
    const double x = ((double)xy[i].x - K.cx) / K.fx, y = ((double)xy[i].y - K.cy) / K.fy;
    const double nn = sqrt(x * x + y * y + 1);
    f[i][0] = x / nn; f[i][1] = y / nn; f[i][2] = 1 / nn;
  }
  double Rs[4][9], ts[4][3];
  const int n = p3p_solve3(P, f, Rs, ts);
  int best = -1;
  double best_e = 0;
  const double X = xyz[3].x, Y = xyz[3].y, Z = xyz[3].z;
  for (int i = 0; i < n; i++) {
    const double xc = Rs[i][0] * X + Rs[i][1] * Y + Rs[i][2] * Z + ts[i][0];
    const double yc = Rs[i][3] * X + Rs[i][4] * Y + Rs[i][5] * Z + ts[i][1];
    const double zc = Rs[i][6] * X + Rs[i][7] * Y + Rs[i][8] * Z + ts[i][2];
    const double u = K.cx + K.fx * xc / zc, v = K.cy + K.fy * yc / zc;
    const double e = (u - xy[3].x) * (u - xy[3].x) + (v - xy[3].y) * (v - xy[3].y);
    if (best < 0 || best_e > e) {
      best = i;
      best_e = e;
    }
  }
  if (best < 0) {
    counts[0] = -1;
    return;
  }
  double rvec[3], R2[9];
  rodrigues_mat2vec(Rs[best], rvec);
  rodrigues_vec2mat(rvec, R2);
  for (int i = 0; i < 3; i++) {
    models[i] = rvec[i];
    models[3 + i] = ts[best][i];
  }
  for (int i = 0; i < 9; i++) models[6 + i] = R2[i];
  models[15] = 0;
  counts[0] = 0;
}

// outputs of the direct (non-RANSAC) solve: every point is an inlier, the pose is the model itself
__global__ void pnp_direct_finish_kernel(const double* __restrict__ models, const int32_t* __restrict__ counts, int n,
                                         int* __restrict__ sel, int32_t* __restrict__ idx, int* __restrict__ n_inl,
                                         double* __restrict__ pose) {
  const int ok = counts[0] >= 0;
  if (threadIdx.x < n) idx[threadIdx.x] = threadIdx.x;
  if (threadIdx.x == 0) {
    sel[0] = ok ? 0 : -1;
    sel[1] = 1;
    sel[2] = ok ? n : 0;
    sel[3] = 1;
    *n_inl = ok ? n : 0;
    for (int i = 0; i < 6; i++) pose[i] = models[i];
  }
}

int pnp_direct_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, int32_t* d_samples, double* d_models,
                      int32_t* d_counts, int* d_sel, int32_t* d_idx, int* d_n_inl, double* d_pose) {
  if (n == 5) {
    const int32_t five[5] = {0, 1, 2, 3, 4};
    VO_CUDA(cudaMemcpyAsync(d_samples, five, sizeof(five), cudaMemcpyHostToDevice, c->stream));
    VO_TRY(pnp_solve_launch(c, xyz, xy, d_samples, 1, d_models, d_counts));
  } else if (n == 4) {
    LaunchScope ls(c, VO_K_PNP_SOLVE);
    pnp_p3p4_kernel<<<1, 32, 0, c->stream>>>(xyz, xy, intr(c), d_models, d_counts);
  } else {
    return VO_ERR_INVALID_ARG;
  }
  {
    LaunchScope ls(c, VO_K_SELECT);
    pnp_direct_finish_kernel<<<1, 32, 0, c->stream>>>(d_models, d_counts, n, d_sel, d_idx, d_n_inl, d_pose);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ================================================================ small point sets
// cv::findFundamentalMat(FM_RANSAC) with 8 <= N <= 14 points does not run RANSAC: OpenCV switches to its LMedS
// estimator (calib3d fundam.cpp: `(method & ~3) == FM_RANSAC && npoints >= 15`, else createLMeDSPointSetRegistrator).
// LMeDSPointSetRegistrator::run: a FIXED number of samples (RANSACUpdateNumIters(conf, 0.45, 7, maxIters), drawn like
// RANSAC's), every model's error vector is sorted, the model with the smallest median wins (first one on ties), then
// sigma = 2.5 * 1.4826 * (1 + 5 / (N - 7)) * sqrt(median), at least 0.001, and the mask is err <= sigma^2.
// One CTA: thread t handles the models t, t + blockDim, ...; errors are OpenCV's (fmat_err, float).
constexpr int LMEDS_THREADS = 256;
constexpr int LMEDS_MAX_N = 14;

__global__ void __launch_bounds__(LMEDS_THREADS)
fmat_lmeds_kernel(const float2* __restrict__ m1, const float2* __restrict__ m2, int n, const double* __restrict__ models,
                  const int32_t* __restrict__ counts, int n_models, int* __restrict__ sel, uint8_t* __restrict__ mask) {
  __shared__ double s_med[LMEDS_THREADS];
  __shared__ int s_idx[LMEDS_THREADS];
  __shared__ float2 s1[LMEDS_MAX_N], s2[LMEDS_MAX_N];
  const int t = threadIdx.x;
  if (t < n) {
    s1[t] = m1[t];
    s2[t] = m2[t];
  }
  __syncthreads();
  double best = 1.7976931348623157e308;   // DBL_MAX: `median < minMedian` never accepts a NaN or an infinite median
  int best_i = -1;
  for (int m = t; m < n_models; m += blockDim.x) {
    if (counts[m] < 0) continue;           // sample with fewer than 3 solutions
    float e[LMEDS_MAX_N];
    for (int i = 0; i < n; i++) e[i] = fmat_err(models + (size_t)m * F_STRIDE, s1[i], s2[i]);
    // OpenCV sorts the float errors as int32 bit patterns (they are non-negative)
    for (int i = 1; i < n; i++) {
      const float v = e[i];
      const int vi = __float_as_int(v);
      int j = i - 1;
      while (j >= 0 && __float_as_int(e[j]) > vi) {
        e[j + 1] = e[j];
        j--;
      }
      e[j + 1] = v;
    }
    const double med = (n & 1) ? (double)e[n / 2] : (double)__fadd_rn(e[n / 2 - 1], e[n / 2]) * 0.5;
    if (med < best) {      // models of one thread are visited in increasing order: first minimum kept
      best = med;
      best_i = m;
    }
  }
  s_med[t] = best;
  s_idx[t] = best_i;
  __syncthreads();
  if (t == 0) {
    double bm = 1.7976931348623157e308;
    int bi = -1;
    for (int k = 0; k < blockDim.x; k++) {
      const int i = s_idx[k];
      if (i < 0) continue;
      if (s_med[k] < bm || (s_med[k] == bm && i < bi)) {   // the sequential loop keeps the FIRST model with the minimum
        bm = s_med[k];
        bi = i;
      }
    }
    int good = 0;
    if (bi >= 0) {
      double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(bm);
      sigma = sigma > 0.001 ? sigma : 0.001;
      const float thr = (float)(sigma * sigma);
      for (int i = 0; i < n; i++) {
        const uint8_t f = fmat_err(models + (size_t)bi * F_STRIDE, s1[i], s2[i]) <= thr;
        mask[i] = f;
        good += f;
      }
      if (good < 7) bi = -1;               // LMeDS: result = count >= modelPoints
    }
    if (bi < 0)
      for (int i = 0; i < n; i++) mask[i] = 0;
    sel[0] = bi;
    sel[1] = n_models / 3;
    sel[2] = good;
    sel[3] = 1;
  }
}

int fmat_lmeds_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, const int32_t* d_counts,
                      int n_models, int* d_sel, uint8_t* d_mask) {
  if (n < 8 || n > LMEDS_MAX_N) return VO_ERR_INVALID_ARG;
  {
    LaunchScope ls(c, VO_K_SELECT);
    fmat_lmeds_kernel<<<1, LMEDS_THREADS, 0, c->stream>>>(m1, m2, n, d_models, d_counts, n_models, d_sel, d_mask);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// N == 7: OpenCV returns the raw 7-point result and a mask of ones (fundam.cpp: `npoints == 7 -> cb->runKernel`)
__global__ void fmat_seven_kernel(const int32_t* __restrict__ counts, int n, int* __restrict__ sel, uint8_t* __restrict__ mask) {
  if (threadIdx.x < n) mask[threadIdx.x] = 1;
  if (threadIdx.x == 0) {
    sel[0] = counts[0] >= 0 ? 0 : -1;
    sel[1] = 1;
    sel[2] = n;
    sel[3] = 1;
  }
}

int fmat_seven_launch(vo_ctx* c, const int32_t* d_counts, int n, int* d_sel, uint8_t* d_mask) {
  {
    LaunchScope ls(c, VO_K_SELECT);
    fmat_seven_kernel<<<1, 32, 0, c->stream>>>(d_counts, n, d_sel, d_mask);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int select_launch(vo_ctx* c, const int32_t* d_counts, int n_samples, int models_per_sample, int model_points,
                  int n_points, double conf, int max_iters, int* d_sel) {
  {
    LaunchScope ls(c, VO_K_SELECT);
    select_kernel<<<1, SEL_THREADS, 0, c->stream>>>(d_counts, n_samples, models_per_sample, model_points, n_points,
                                                    conf, max_iters, d_sel, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
