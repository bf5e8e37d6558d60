// cvmath.cuh -- FP64 small-matrix numerics of the VO hot path, written so that the CUDA
// kernels reproduce OpenCV 4.13's results operation for operation.
//
// Why op-for-op: the reference (Gautham-JS/ROS_Stereo_SLAM) calls cv::solvePnPRansac
// (src/keyFrameManagement.cpp:84), cv::findFundamentalMat (src/tracking.cpp:34,75) and
// cv::triangulatePoints (src/triangulation.cpp:152).  Their minimal solvers take null
// spaces of rank-deficient matrices with OpenCV's one-sided Jacobi SVD; the basis that
// comes out is decided by rounding, so "inlier sets bit-exact for the same sample list"
// needs the same rounding.  Everything here uses only + - * / sqrt on IEEE doubles in a
// fixed order (compile with --fmad=false), which an sm_100a FP64 pipe and an x86 SSE2
// build execute identically.  OpenCV's source is not in the reference tree; these are
// restatements of its published algorithms, pinned against cv2 4.13.0 by
// tests/test_hostmath.py (host build of this very header) and on the GPU by
// tests/test_gpu_*.py.
//
// All functions are `static inline` host+device; no global state, no allocation.
#pragma once
#include <math.h>
#include <float.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VO_HD __host__ __device__ __forceinline__
#define VO_HDN static __host__ __device__ __noinline__
#define VO_HDM __host__ __device__ __forceinline__
#ifdef VO_HELPERS_NOINLINE
#define VO_HDF static __host__ __device__ __noinline__
#else
#define VO_HDF __host__ __device__ __forceinline__
#endif
#else
#define VO_HDF static inline
#define VO_HD static inline
#define VO_HDN static
#define VO_HDM inline
#endif

#ifndef VO_JACOBI_SPEC_M3
#define VO_JACOBI_SPEC_M3 1
#endif

namespace vo {

// ---------------------------------------------------------------------------------------
// OpenCV's own hypot (lapack.cpp shadows ::hypot inside namespace cv).
// One instruction stream for both orderings (the lanes of a warp that rotate different row pairs do not diverge):
// the larger magnitude divides the smaller one, exactly as the two branches of OpenCV's code do.
VO_HD double cv_hypot(double a, double b) {
  a = fabs(a);
  b = fabs(b);
  const bool a_larger = a > b;
  const double hi = a_larger ? a : b, lo = a_larger ? b : a;
  if (!(hi > 0)) return 0;
  const double q = lo / hi;
  return hi * sqrt(1 + q * q);
}

// The rotation of OpenCV's one-sided Jacobi (JacobiSVDImpl_): beta < 0: s = sqrt(delta / gamma), c = p / (gamma * s * 2);
// otherwise c = sqrt((gamma + beta) / (gamma * 2)), s = p / (gamma * c * 2) -- the same two operations on selected
// operands, written once so that lanes with different signs of beta share the instruction stream.
VO_HD void cv_jacobi_cs(double p, double beta, double gamma, double& c, double& s) {
  const bool neg = beta < 0;
  const double num = neg ? (gamma - beta) * 0.5 : gamma + beta;
  const double den = neg ? gamma : gamma * 2;
  const double r1 = sqrt(num / den);
  const double r2 = p / (gamma * r1 * 2);
  s = neg ? r1 : r2;
  c = neg ? r2 : r1;
}

// cv::RNG (multiply-with-carry)
struct CvRng {
  uint64_t state;
  VO_HDM explicit CvRng(uint64_t s = 0xffffffffffffffffULL) : state(s ? s : 0xffffffffULL) {}
  VO_HDM uint32_t next() {
    state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
    return (uint32_t)state;
  }
  VO_HDM int uniform(int a, int b) { return a + (int)(next() % (uint32_t)(b - a)); }
};

// ---------------------------------------------------------------------------------------
// JacobiSVDImpl_<double> (one-sided Jacobi / Hestenes).  At: n rows of length m (row
// stride astep), orthogonalised in place; on return row i = sigma_i * u_i normalised to
// u_i.  Vt: n x n (stride vstep), W: n singular values, descending.  Rows n..n1-1 of At
// (zero singular values when n1 > n, or exactly-zero ones) are completed with OpenCV's
// deterministic pseudo-random Gram-Schmidt vectors (RNG 0x12345678).
// Compile-time shapes (M = row length, N = rows, N1 = rows to complete) let the compiler
// unroll the length-M inner loops and keep a row pair in registers across dot product,
// rotation and norm update; the arithmetic and its order are unchanged.
// WITH_V=false skips the accumulation of V (its rotations never feed back into At/W, so U
// and W are bit-identical); callers that only consume U^T (EPnP's 12x12 and 3x3 PCA) use it
// and pass any non-null pointer as Vt (OpenCV's code keys the row normalisation on Vt != 0).
// SKIP_SWEEPS=true: the rows of At were already orthogonalised (by jacobi_sweeps_warp on the
// device); only the norms, the ordering and the completion of the rows are done here.
template <int M, int N, int N1, bool WITH_V = true, bool SKIP_SWEEPS = false>
VO_HDN void jacobi_svd(double* At, double* _W, double* Vt) {
  constexpr int astep = M, vstep = N, m = M, n = N, n1 = N1;
  double W[N];
  const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
  int i, j, k, iter;
  constexpr int max_iter = m > 30 ? m : 30;
  double c, s, sd;

  for (i = 0; i < n; i++) {
    sd = 0;
#pragma unroll
    for (k = 0; k < m; k++) {
      double t = At[i * astep + k];
      sd += t * t;
    }
    W[i] = sd;
    if (WITH_V && Vt) {
#pragma unroll
      for (k = 0; k < n; k++) Vt[i * vstep + k] = 0;
      Vt[i * vstep + i] = 1;
    }
  }

  for (iter = 0; iter < (SKIP_SWEEPS ? 0 : max_iter); iter++) {
    bool changed = false;
    for (i = 0; i < n - 1; i++)
      for (j = i + 1; j < n; j++) {
        double *Ai = At + i * astep, *Aj = At + j * astep;
        double a = W[i], p = 0, b = W[j];
        double ri[M], rj[M];
#pragma unroll
        for (k = 0; k < m; k++) {
          ri[k] = Ai[k];
          rj[k] = Aj[k];
        }
#pragma unroll
        for (k = 0; k < m; k++) p += ri[k] * rj[k];
        // latency-bound callers (SPEC): the skip test's square root and the hypot of the rotation are independent
        // chains, started together; throughput-bound ones (triangulation) keep the test first
        constexpr bool SPEC = VO_JACOBI_SPEC_M3 && M == 3;
        double beta = a - b, gamma = 0;
        const double p2 = p * 2;
        if (SPEC) gamma = cv_hypot(p2, beta);
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p = p2;
        if (!SPEC) gamma = cv_hypot(p, beta);
        cv_jacobi_cs(p, beta, gamma, c, s);

        a = b = 0;
#pragma unroll
        for (k = 0; k < m; k++) {
          double t0 = c * ri[k] + s * rj[k];
          double t1 = -s * ri[k] + c * rj[k];
          Ai[k] = t0;
          Aj[k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        W[i] = a;
        W[j] = b;
        changed = true;

        if (WITH_V && Vt) {
          double *Vi = Vt + i * vstep, *Vj = Vt + j * vstep;
#pragma unroll
          for (k = 0; k < n; k++) {
            double t0 = c * Vi[k] + s * Vj[k];
            double t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0;
            Vj[k] = t1;
          }
        }
      }
    if (!changed) break;
  }

  for (i = 0; i < n; i++) {
    sd = 0;
#pragma unroll
    for (k = 0; k < m; k++) {
      double t = At[i * astep + k];
      sd += t * t;
    }
    W[i] = sqrt(sd);
  }

  for (i = 0; i < n - 1; i++) {
    j = i;
    for (k = i + 1; k < n; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i];
      W[i] = W[j];
      W[j] = t;
      if (Vt) {
        for (k = 0; k < m; k++) {
          t = At[i * astep + k];
          At[i * astep + k] = At[j * astep + k];
          At[j * astep + k] = t;
        }
        if (WITH_V)
          for (k = 0; k < n; k++) {
            t = Vt[i * vstep + k];
            Vt[i * vstep + k] = Vt[j * vstep + k];
            Vt[j * vstep + k] = t;
          }
      }
    }
  }

  for (i = 0; i < n; i++) _W[i] = W[i];
  if (!Vt) return;

  CvRng rng(0x12345678);
  for (i = 0; i < n1; i++) {
    sd = i < n ? W[i] : 0;
    for (int ii = 0; ii < 100 && sd <= minval; ii++) {
      // zero singular value: random +-1/m vector, two Gram-Schmidt passes against the
      // previous rows, L1-normalised after each projection (as OpenCV does)
      const double val0 = 1. / m;
      for (k = 0; k < m; k++) {
        double val = (rng.next() & 256) != 0 ? val0 : -val0;
        At[i * astep + k] = val;
      }
      for (iter = 0; iter < 2; iter++) {
        for (j = 0; j < i; j++) {
          sd = 0;
          for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
          double asum = 0;
          for (k = 0; k < m; k++) {
            double t = At[i * astep + k] - sd * At[j * astep + k];
            At[i * astep + k] = t;
            asum += fabs(t);
          }
          asum = asum > eps * 100 ? 1 / asum : 0;
          for (k = 0; k < m; k++) At[i * astep + k] *= asum;
        }
      }
      sd = 0;
      for (k = 0; k < m; k++) {
        double t = At[i * astep + k];
        sd += t * t;
      }
      sd = sqrt(sd);
    }
    s = sd > minval ? 1 / sd : 0.;
    for (k = 0; k < m; k++) At[i * astep + k] *= s;
  }
}

// cv::SVD::compute(A) for a SQUARE N x N matrix A (row-major): w (N), u (N x N, columns
// = left vectors), vt (N x N).  Mirrors _SVDcompute: temp_a = A^T, Jacobi on its rows.
template <int N>
VO_HDF void svd_square(const double* A, double* w, double* u, double* vt) {
  double at[N * N];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) at[i * N + j] = A[j * N + i];
  jacobi_svd<N, N, N>(at, w, vt);
  if (u)
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) u[i * N + j] = at[j * N + i];
}

// Same, but returns U^T (rows = left singular vectors) -- the CV_SVD_U_T form EPnP uses.
template <int N>
VO_HDF void svd_square_ut(const double* A, double* w, double* ut, double* vt) {
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) ut[i * N + j] = A[j * N + i];
  jacobi_svd<N, N, N>(ut, w, vt);
}

// U^T and W only (V not accumulated; `ut` doubles as the non-null marker OpenCV's code needs
// to run its normalisation / zero-singular-value completion of the rows).
template <int N>
VO_HDF void svd_square_ut_only(const double* A, double* w, double* ut) {
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) ut[i * N + j] = A[j * N + i];
  jacobi_svd<N, N, N, false>(ut, w, ut);
}

// cv::solve(A (M x N, M >= N), b (M), x (N), DECOMP_SVD): Jacobi SVD of A^T's rows + SVBkSb.
template <int M, int N>
VO_HDF void solve_svd(const double* A, const double* b, double* x) {
  double at[N * M], w[N], v[N * N];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < M; j++) at[i * M + j] = A[j * N + i];
  jacobi_svd<M, N, N>(at, w, v);
  // SVBkSbImpl_ with nb == 1, u = at (uT), v (vT), eps = 2*DBL_EPSILON
  double threshold = 0;
  for (int i = 0; i < N; i++) {
    x[i] = 0;
    threshold += w[i];
  }
  threshold *= DBL_EPSILON * 2;
  for (int i = 0; i < N; i++) {
    double wi = w[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double s = 0;
    for (int j = 0; j < M; j++) s += at[i * M + j] * b[j];
    s *= wi;
    for (int j = 0; j < N; j++) x[j] = x[j] + s * v[i * N + j];
  }
}

// Runtime-size variant of jacobi_svd (n1 = n, V accumulated) and of cv::solve(DECOMP_SVD): same
// operations in the same order as the compile-time versions.  EPnP's three beta initialisations
// solve 6x4, 6x3 and 6x5 systems; on the device they run on three lanes of one warp, and sharing
// ONE instruction stream (instead of three template instantiations) keeps those lanes converged.
// The three parts are separate functions because the device runs the sweeps of EPnP's three 6 x n problems as
// wavefronts over lane groups (jacobi_warp.cuh, jacobi_sweeps_groups) around the same init / rotate / finish code.
// Vt: n rows of stride NMAX (columns n..NMAX-1 stay zero).
template <int M, int NMAX>
VO_HD void jacobi_rt_init(const double* At, double* W, double* Vt, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    double sd = 0;
#pragma unroll
    for (int k = 0; k < M; k++) {
      double t = At[i * M + k];
      sd += t * t;
    }
    W[i] = sd;
#pragma unroll
    for (int k = 0; k < NMAX; k++) Vt[i * NMAX + k] = 0;
    Vt[i * NMAX + i] = 1;
  }
}

// one pair (i, j) of a sweep: rows i and j of At and Vt, W[i], W[j]; returns whether it rotated
template <int M, int NMAX>
VO_HD bool jacobi_rt_rotate(double* At, double* W, double* Vt, int i, int j) {
  const double eps = DBL_EPSILON * 10;
  double *Ai = At + i * M, *Aj = At + j * M;
  double a = W[i], p = 0, b = W[j];
  double ri[M], rj[M], vi[NMAX], vj[NMAX];
  double *Vi = Vt + i * NMAX, *Vj = Vt + j * NMAX;
#pragma unroll
  for (int k = 0; k < M; k++) {
    ri[k] = Ai[k];
    rj[k] = Aj[k];
  }
#pragma unroll
  for (int k = 0; k < NMAX; k++) {   // requested early: the loads complete under the divide / square-root chain
    vi[k] = Vi[k];
    vj[k] = Vj[k];
  }
#pragma unroll
  for (int k = 0; k < M; k++) p += ri[k] * rj[k];
  const double skip_thr = eps * sqrt(a * b);
  const double p2 = p * 2;
  double beta = a - b, gamma = cv_hypot(p2, beta);   // evaluated next to the skip test's square root (two independent chains)
  if (fabs(p) <= skip_thr) return false;
  p = p2;
  double c, s;
  cv_jacobi_cs(p, beta, gamma, c, s);
  a = b = 0;
#pragma unroll
  for (int k = 0; k < M; k++) {
    double t0 = c * ri[k] + s * rj[k];
    double t1 = -s * ri[k] + c * rj[k];
    Ai[k] = t0;
    Aj[k] = t1;
    a += t0 * t0;
    b += t1 * t1;
  }
  W[i] = a;
  W[j] = b;
#pragma unroll
  for (int k = 0; k < NMAX; k++) {
    double t0 = c * vi[k] + s * vj[k];
    double t1 = -s * vi[k] + c * vj[k];
    Vi[k] = t0;
    Vj[k] = t1;
  }
  return true;
}

// after the sweeps: singular values, ordering, normalisation of the rows (completion of zero rows)
template <int M, int NMAX>
VO_HD void jacobi_rt_finish(double* At, double* _W, double* Vt, int n) {
  constexpr int astep = M, m = M, vstep = NMAX;
  double W[8];
  const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
  int i, j, k, iter;
  double s, sd;
#pragma unroll 1
  for (i = 0; i < n; i++) {
    sd = 0;
#pragma unroll
    for (k = 0; k < m; k++) {
      double t = At[i * astep + k];
      sd += t * t;
    }
    W[i] = sqrt(sd);
  }
#pragma unroll 1
  for (i = 0; i < n - 1; i++) {
    j = i;
#pragma unroll 1
    for (k = i + 1; k < n; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i];
      W[i] = W[j];
      W[j] = t;
#pragma unroll 1
      for (k = 0; k < m; k++) {
        t = At[i * astep + k];
        At[i * astep + k] = At[j * astep + k];
        At[j * astep + k] = t;
      }
#pragma unroll 1
      for (k = 0; k < n; k++) {
        t = Vt[i * vstep + k];
        Vt[i * vstep + k] = Vt[j * vstep + k];
        Vt[j * vstep + k] = t;
      }
    }
  }
#pragma unroll 1
  for (i = 0; i < n; i++) _W[i] = W[i];
  CvRng rng(0x12345678);
#pragma unroll 1
  for (i = 0; i < n; i++) {
    sd = W[i];
#pragma unroll 1
    for (int ii = 0; ii < 100 && sd <= minval; ii++) {
      const double val0 = 1. / m;
#pragma unroll 1
      for (k = 0; k < m; k++) {
        double val = (rng.next() & 256) != 0 ? val0 : -val0;
        At[i * astep + k] = val;
      }
#pragma unroll 1
      for (iter = 0; iter < 2; iter++) {
#pragma unroll 1
        for (j = 0; j < i; j++) {
          sd = 0;
#pragma unroll 1
          for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
          double asum = 0;
#pragma unroll 1
          for (k = 0; k < m; k++) {
            double t = At[i * astep + k] - sd * At[j * astep + k];
            At[i * astep + k] = t;
            asum += fabs(t);
          }
          asum = asum > eps * 100 ? 1 / asum : 0;
#pragma unroll 1
          for (k = 0; k < m; k++) At[i * astep + k] *= asum;
        }
      }
      sd = 0;
#pragma unroll 1
      for (k = 0; k < m; k++) {
        double t = At[i * astep + k];
        sd += t * t;
      }
      sd = sqrt(sd);
    }
    s = sd > minval ? 1 / sd : 0.;
#pragma unroll
    for (k = 0; k < m; k++) At[i * astep + k] *= s;
  }
}

template <int M, int NMAX>
VO_HDN void jacobi_svd_rt(double* At, double* _W, double* Vt, int n) {
  double W[8];
  constexpr int max_iter = M > 30 ? M : 30;
  jacobi_rt_init<M, NMAX>(At, W, Vt, n);
#pragma unroll 1
  for (int iter = 0; iter < max_iter; iter++) {
    bool changed = false;
#pragma unroll 1
    for (int i = 0; i < n - 1; i++)
#pragma unroll 1
      for (int j = i + 1; j < n; j++) changed |= jacobi_rt_rotate<M, NMAX>(At, W, Vt, i, j);
    if (!changed) break;
  }
  jacobi_rt_finish<M, NMAX>(At, _W, Vt, n);
}

// cv::solve(A (6 x n), b (6), x (n), DECOMP_SVD) for n <= 5
// SVBkSbImpl_ with nb == 1 on the Jacobi result: u = at (uT, n rows of 6), v (vT, stride 5), eps = 2 * DBL_EPSILON
VO_HD void svd_backsubst_6xn(const double* at, const double* w, const double* v, const double* b, double* x, int n) {
  const int m = 6;
  double threshold = 0;
  for (int i = 0; i < n; i++) {
    x[i] = 0;
    threshold += w[i];
  }
  threshold *= DBL_EPSILON * 2;
  for (int i = 0; i < n; i++) {
    double wi = w[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double s = 0;
    for (int j = 0; j < m; j++) s += at[i * m + j] * b[j];
    s *= wi;
    for (int j = 0; j < n; j++) x[j] = x[j] + s * v[i * 5 + j];
  }
}

// cv::invert(A 3x3, DECOMP_SVD) = SVD::compute + SVD::backSubst(w, u, vt, Mat(), dst).
VO_HDF void invert3_svd(const double* A, double* inv) {
  double w[3], u[9], vt[9];
  svd_square<3>(A, w, u, vt);
  double threshold = 0;
  for (int i = 0; i < 9; i++) inv[i] = 0;
  for (int i = 0; i < 3; i++) threshold += w[i];
  threshold *= DBL_EPSILON * 2;
  for (int i = 0; i < 3; i++) {
    double wi = w[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double buffer[3];
    for (int j = 0; j < 3; j++) buffer[j] = u[j * 3 + i] * wi;  // u not transposed: u[j*ldu + i]
    for (int r = 0; r < 3; r++) {                               // MatrAXPY(n, nb, buffer, 0, v, vdelta1, x, ldx)
      double sv = vt[i * 3 + r];
      for (int j = 0; j < 3; j++) inv[r * 3 + j] = inv[r * 3 + j] + sv * buffer[j];
    }
  }
}

// cv::mulTransposed(src (rows x COLS), dst, aTa=true): dst = src^T src, upper triangle
// accumulated sequentially over rows, then mirrored (completeSymm).
// FMA=true reproduces the AVX2/AVX-512 dispatch of OpenCV's matmul kernel, where the
// compiler contracts `s += a*b`.
template <int COLS, bool FMA>
VO_HDF void mul_transposed(const double* src, int rows, double* dst) {
  for (int i = 0; i < COLS; i++) {
    for (int j = i; j < COLS; j++) {
      double s0 = 0;
      for (int k = 0; k < rows; k++) {
        if (FMA)
          s0 = fma(src[k * COLS + i], src[k * COLS + j], s0);
        else
          s0 += src[k * COLS + i] * src[k * COLS + j];
      }
      dst[i * COLS + j] = s0;
    }
  }
  for (int i = 0; i < COLS; i++)
    for (int j = 0; j < i; j++) dst[i * COLS + j] = dst[j * COLS + i];
}

// ---------------------------------------------------------------------------------------
// cv::Rodrigues, both directions (smooth; ulp-level libm differences are harmless).
VO_HDF void rodrigues_vec2mat(const double r_[3], double R[9]) {
  double rx = r_[0], ry = r_[1], rz = r_[2];
  double theta = sqrt(rx * rx + ry * ry + rz * rz);
  if (theta < DBL_EPSILON) {
    R[0] = 1; R[1] = 0; R[2] = 0;
    R[3] = 0; R[4] = 1; R[5] = 0;
    R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  double c = cos(theta), s = sin(theta), c1 = 1. - c;
  double itheta = theta ? 1. / theta : 0.;
  rx *= itheta; ry *= itheta; rz *= itheta;
  R[0] = c + c1 * rx * rx;      R[1] = c1 * rx * ry - s * rz; R[2] = c1 * rx * rz + s * ry;
  R[3] = c1 * rx * ry + s * rz; R[4] = c + c1 * ry * ry;      R[5] = c1 * ry * rz - s * rx;
  R[6] = c1 * rx * rz - s * ry; R[7] = c1 * ry * rz + s * rx; R[8] = c + c1 * rz * rz;
}

VO_HDF void rodrigues_mat2vec(const double Rin[9], double r[3]) {
  double w[3], u[9], vt[9], R[9];
  svd_square<3>(Rin, w, u, vt);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 3 + j] = u[i * 3 + 0] * vt[0 * 3 + j] + u[i * 3 + 1] * vt[1 * 3 + j] + u[i * 3 + 2] * vt[2 * 3 + j];
  double x = R[7] - R[5], y = R[2] - R[6], z = R[3] - R[1];
  double s = sqrt((x * x + y * y + z * z) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      x = y = z = 0;
    } else {
      double t = (R[0] + 1) * 0.5;
      x = sqrt(t > 0. ? t : 0.);
      t = (R[4] + 1) * 0.5;
      y = sqrt(t > 0. ? t : 0.) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5;
      z = sqrt(t > 0. ? t : 0.) * (R[2] < 0 ? -1. : 1.);
      if (fabs(x) < fabs(y) && fabs(x) < fabs(z) && (R[5] > 0) != (y * z > 0)) z = -z;
      theta /= sqrt(x * x + y * y + z * z);
      x *= theta; y *= theta; z *= theta;
    }
  } else {
    double vth = 1 / (2 * s);
    vth *= theta;
    x *= vth; y *= vth; z *= vth;
  }
  r[0] = x; r[1] = y; r[2] = z;
}

// ---------------------------------------------------------------------------------------
// EPnP on exactly 5 correspondences = the minimal solver inside cv::solvePnPRansac
// (PnPRansacCallback::runKernel -> solvePnP(SOLVEPNP_EPNP)).  Follows OpenCV's epnp.cpp
// function by function (choose_control_points, compute_barycentric_coordinates, fill_M,
// compute_L_6x10, compute_rho, find_betas_approx_{1,2,3}, gauss_newton/qr_solve,
// compute_R_and_t).  Inputs: 5 object points (float32, as the RANSAC subset holds them)
// and 5 image points (float32 pixels); output rvec/tvec like the RANSAC model column.
struct Intrinsics {
  double fx, fy, cx, cy;
};

struct EpnpWork {   // read-only after epnp5_front
  double pws[15], us[10], alphas[20];
  double cws[4][3];
};

VO_HD double epnp_dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
VO_HD double epnp_dist2(const double* p1, const double* p2) {
  return (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) + (p1[2] - p2[2]) * (p1[2] - p2[2]);
}

VO_HDF void epnp_qr_solve_6x4(double* pA, double* pb, double* pX) {
  const int nr = 6, nc = 4;
  double A1[6], A2[6];
  double* ppAkk = pA;
  for (int k = 0; k < nc; k++) {
    double* ppAik1 = ppAkk;
    double eta = fabs(*ppAik1);
    for (int i = k + 1; i < nr; i++) {
      double elt = fabs(*ppAik1);
      if (eta < elt) eta = elt;
      ppAik1 += nc;
    }
    if (eta == 0) {
      A1[k] = A2[k] = 0.0;
      return;
    } else {
      double* ppAik2 = ppAkk;
      double sum2 = 0.0, inv_eta = 1. / eta;
      for (int i = k; i < nr; i++) {
        *ppAik2 *= inv_eta;
        sum2 += *ppAik2 * *ppAik2;
        ppAik2 += nc;
      }
      double sigma = sqrt(sum2);
      if (*ppAkk < 0) sigma = -sigma;
      *ppAkk += sigma;
      A1[k] = sigma * *ppAkk;
      A2[k] = -eta * sigma;
      for (int j = k + 1; j < nc; j++) {
        double* ppAik = ppAkk;
        double sum = 0;
        for (int i = k; i < nr; i++) {
          sum += *ppAik * ppAik[j - k];
          ppAik += nc;
        }
        double tau = sum / A1[k];
        ppAik = ppAkk;
        for (int i = k; i < nr; i++) {
          ppAik[j - k] -= tau * *ppAik;
          ppAik += nc;
        }
      }
    }
    ppAkk += nc + 1;
  }
  // b <- Qt b
  double* ppAjj = pA;
  for (int j = 0; j < nc; j++) {
    double* ppAij = ppAjj;
    double tau = 0;
    for (int i = j; i < nr; i++) {
      tau += *ppAij * pb[i];
      ppAij += nc;
    }
    tau /= A1[j];
    ppAij = ppAjj;
    for (int i = j; i < nr; i++) {
      pb[i] -= tau * *ppAij;
      ppAij += nc;
    }
    ppAjj += nc + 1;
  }
  // X = R^-1 b
  pX[nc - 1] = pb[nc - 1] / A2[nc - 1];
  for (int i = nc - 2; i >= 0; i--) {
    double* ppAij = pA + i * nc + (i + 1);
    double sum = 0;
    for (int j = i + 1; j < nc; j++) {
      sum += *ppAij * pX[j];
      ppAij++;
    }
    pX[i] = (pb[i] - sum) / A2[i];
  }
}

VO_HDF void epnp_gauss_newton(const double* l_6x10, const double* rho, double betas[4]) {
  for (int it = 0; it < 5; it++) {
    double a[24], b[6], x[4] = {0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
      const double* rowL = l_6x10 + i * 10;
      double* rowA = a + i * 4;
      rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
      rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
      rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
      rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
      b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                       rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                       rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                       rowL[9] * betas[3] * betas[3]);
    }
    epnp_qr_solve_6x4(a, b, x);
    for (int i = 0; i < 4; i++) betas[i] += x[i];
  }
}

VO_HDF void epnp_estimate_R_and_t(const EpnpWork& w, const double* pcs, double R[3][3], double t[3]) {
  const int n = 5;
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++) {
      pc0[j] += pcs[3 * i + j];
      pw0[j] += w.pws[3 * i + j];
    }
  for (int j = 0; j < 3; j++) {
    pc0[j] /= n;
    pw0[j] /= n;
  }
  double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, abt_d[3], abt_u[9], abt_vt[9];
  for (int i = 0; i < n; i++) {
    const double* pc = &pcs[3 * i];
    const double* pw = &w.pws[3 * i];
    for (int j = 0; j < 3; j++) {
      abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
      abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
      abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
    }
  }
  // cvSVD(&ABt, &D, &U, &V, CV_SVD_MODIFY_A): U, V (not transposed); R = U * V^T
  svd_square<3>(abt, abt_d, abt_u, abt_vt);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      // dot(abt_u row i, abt_v row j) with abt_v = vt^T  ->  sum_k u[i][k] * vt[k][j]
      R[i][j] = abt_u[3 * i] * abt_vt[j] + abt_u[3 * i + 1] * abt_vt[3 + j] + abt_u[3 * i + 2] * abt_vt[6 + j];
    }
  const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                     R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
  if (det < 0) {
    R[2][0] = -R[2][0];
    R[2][1] = -R[2][1];
    R[2][2] = -R[2][2];
  }
  t[0] = pc0[0] - epnp_dot3(R[0], pw0);
  t[1] = pc0[1] - epnp_dot3(R[1], pw0);
  t[2] = pc0[2] - epnp_dot3(R[2], pw0);
}

VO_HDF double epnp_compute_R_and_t(const EpnpWork& w, const Intrinsics& K, const double* ut, const double* betas,
                                  double R[3][3], double t[3]) {
  const int n = 5;
  double ccs[4][3], pcs[15];
  // compute_ccs
  for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
  for (int i = 0; i < 4; i++) {
    const double* v = ut + 12 * (11 - i);
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
  }
  // compute_pcs
  for (int i = 0; i < n; i++) {
    const double* a = &w.alphas[4 * i];
    double* pc = &pcs[3 * i];
    for (int j = 0; j < 3; j++) pc[j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
  }
  // solve_for_sign
  if (pcs[2] < 0.0) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 3; j++) ccs[i][j] = -ccs[i][j];
    for (int i = 0; i < n; i++) {
      pcs[3 * i] = -pcs[3 * i];
      pcs[3 * i + 1] = -pcs[3 * i + 1];
      pcs[3 * i + 2] = -pcs[3 * i + 2];
    }
  }
  epnp_estimate_R_and_t(w, pcs, R, t);
  // reprojection_error
  double sum2 = 0.0;
  for (int i = 0; i < n; i++) {
    const double* pw = &w.pws[3 * i];
    double Xc = epnp_dot3(R[0], pw) + t[0];
    double Yc = epnp_dot3(R[1], pw) + t[1];
    double inv_Zc = 1.0 / (epnp_dot3(R[2], pw) + t[2]);
    double ue = K.cx + K.fx * Xc * inv_Zc;
    double ve = K.cy + K.fy * Yc * inv_Zc;
    double u = w.us[2 * i], v = w.us[2 * i + 1];
    sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
  }
  return sum2 / n;
}

// EPnP is split at its 12x12 SVD so that the device can run the Jacobi sweeps of that SVD
// warp-cooperatively (jacobi_warp.cuh) while everything else stays on one lane.
// FMA_MTM: see mul_transposed.
// epnp5_front_M: everything up to the 10x12 matrix M;  MtM = M^T M follows (mul_transposed).
template <bool FMA_MTM>
VO_HDN void epnp5_front_M(const float* obj /*15*/, const float* img /*10*/, const Intrinsics& K, EpnpWork& w,
                          double* M /*120*/) {
  const int n = 5;
  // solvePnP: undistortPoints (float in, float out, zero distortion) then epnp::init_points
  const double ifx = 1. / K.fx, ify = 1. / K.fy;
  for (int i = 0; i < n; i++) {
    w.pws[3 * i] = obj[3 * i];
    w.pws[3 * i + 1] = obj[3 * i + 1];
    w.pws[3 * i + 2] = obj[3 * i + 2];
    float xn = (float)(((double)img[2 * i] - K.cx) * ifx);
    float yn = (float)(((double)img[2 * i + 1] - K.cy) * ify);
    w.us[2 * i] = xn * K.fx + K.cx;
    w.us[2 * i + 1] = yn * K.fy + K.cy;
  }
  // choose_control_points
  w.cws[0][0] = w.cws[0][1] = w.cws[0][2] = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++) w.cws[0][j] += w.pws[3 * i + j];
  for (int j = 0; j < 3; j++) w.cws[0][j] /= n;
  {
    double pw0[15], pw0tpw0[9], dc[3], uct[9];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < 3; j++) pw0[3 * i + j] = w.pws[3 * i + j] - w.cws[0][j];
    mul_transposed<3, FMA_MTM>(pw0, n, pw0tpw0);
    svd_square_ut_only<3>(pw0tpw0, dc, uct);
    for (int i = 1; i < 4; i++) {
      double k = sqrt(dc[i - 1] / n);
      for (int j = 0; j < 3; j++) w.cws[i][j] = w.cws[0][j] + k * uct[3 * (i - 1) + j];
    }
  }
  // compute_barycentric_coordinates
  {
    double cc[9], cc_inv[9];
    for (int i = 0; i < 3; i++)
      for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = w.cws[j][i] - w.cws[0][i];
    invert3_svd(cc, cc_inv);
    const double* ci = cc_inv;
    for (int i = 0; i < n; i++) {
      const double* pi = &w.pws[3 * i];
      double* a = &w.alphas[4 * i];
      for (int j = 0; j < 3; j++)
        a[1 + j] = ci[3 * j] * (pi[0] - w.cws[0][0]) + ci[3 * j + 1] * (pi[1] - w.cws[0][1]) + ci[3 * j + 2] * (pi[2] - w.cws[0][2]);
      a[0] = 1.0f - a[1] - a[2] - a[3];
    }
  }
  // M (10 x 12)
  {
    for (int i = 0; i < n; i++) {
      const double* as = &w.alphas[4 * i];
      double u = w.us[2 * i], v = w.us[2 * i + 1];
      double* M1 = M + (2 * i) * 12;
      double* M2 = M1 + 12;
      for (int q = 0; q < 4; q++) {
        M1[3 * q] = as[q] * K.fx;
        M1[3 * q + 1] = 0.0;
        M1[3 * q + 2] = as[q] * (K.cx - u);
        M2[3 * q] = 0.0;
        M2[3 * q + 1] = as[q] * K.fy;
        M2[3 * q + 2] = as[q] * (K.cy - v);
      }
    }
  }
}

template <bool FMA_MTM>
VO_HDN void epnp5_front(const float* obj /*15*/, const float* img /*10*/, const Intrinsics& K, EpnpWork& w,
                        double* mtm /*144*/) {
  double M[120];
  epnp5_front_M<FMA_MTM>(obj, img, K, w, M);
  mul_transposed<12, FMA_MTM>(M, 10, mtm);
}

// ---- back end: L/rho from the last four left singular vectors, then the three beta
// initialisations N = 1, 2, 3 (each followed by 5 Gauss-Newton steps and a Horn alignment) and
// the choice of the smallest reprojection error.  The three variants are independent: the
// device runs them on three lanes (ransac.cu), the host one after the other.
// ut: U^T of MtM (rows = left singular vectors, descending singular values)
VO_HDN void epnp_prepare(const EpnpWork& w, const double* ut, double* l_6x10 /*60*/, double* rho /*6*/) {
  // compute_L_6x10, compute_rho
  {
    const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
    double dv[4][6][3];
    for (int i = 0; i < 4; i++) {
      int a = 0, b = 1;
      for (int j = 0; j < 6; j++) {
        dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
        dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
        dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
        b++;
        if (b > 3) {
          a++;
          b = a + 1;
        }
      }
    }
    for (int i = 0; i < 6; i++) {
      double* row = l_6x10 + 10 * i;
      row[0] = epnp_dot3(dv[0][i], dv[0][i]);
      row[1] = 2.0f * epnp_dot3(dv[0][i], dv[1][i]);
      row[2] = epnp_dot3(dv[1][i], dv[1][i]);
      row[3] = 2.0f * epnp_dot3(dv[0][i], dv[2][i]);
      row[4] = 2.0f * epnp_dot3(dv[1][i], dv[2][i]);
      row[5] = epnp_dot3(dv[2][i], dv[2][i]);
      row[6] = 2.0f * epnp_dot3(dv[0][i], dv[3][i]);
      row[7] = 2.0f * epnp_dot3(dv[1][i], dv[3][i]);
      row[8] = 2.0f * epnp_dot3(dv[2][i], dv[3][i]);
      row[9] = epnp_dot3(dv[3][i], dv[3][i]);
    }
    rho[0] = epnp_dist2(w.cws[0], w.cws[1]);
    rho[1] = epnp_dist2(w.cws[0], w.cws[2]);
    rho[2] = epnp_dist2(w.cws[0], w.cws[3]);
    rho[3] = epnp_dist2(w.cws[1], w.cws[2]);
    rho[4] = epnp_dist2(w.cws[1], w.cws[3]);
    rho[5] = epnp_dist2(w.cws[2], w.cws[3]);
  }

}

// variant N in {1,2,3}: betas (4, out), R, t; returns the mean reprojection error.
// find_betas_approx_1 uses columns [0 1 3 6] of L, _2 columns [0 1 2], _3 columns [0 1 2 3 4].
// Three steps, so that the device can run the Jacobi sweeps of the three 6 x n systems side by side as wavefronts:
// epnp_variant_system (the transposed system, n rows of 6), the SVD solve, epnp_variant_back (everything after it).
VO_HD int epnp_variant_ncol(int N) { return N == 1 ? 4 : (N == 2 ? 3 : 5); }

VO_HD void epnp_variant_system(int N, const double* l_6x10, double* at /* ncol x 6 */) {
  const int ncol = epnp_variant_ncol(N);
  for (int j = 0; j < ncol; j++) {
    const int col = N == 1 ? (j == 0 ? 0 : j == 1 ? 1 : j == 2 ? 3 : 6) : j;
    for (int i = 0; i < 6; i++) at[j * 6 + i] = l_6x10[i * 10 + col];
  }
}

VO_HDN double epnp_variant_back(int N, const EpnpWork& w, const Intrinsics& K, const double* ut, const double* l_6x10,
                                const double* rho, const double* bx, double* betas, double R[3][3], double t[3]) {
  if (N == 1) {
    if (bx[0] < 0) {
      betas[0] = sqrt(-bx[0]);
      betas[1] = -bx[1] / betas[0];
      betas[2] = -bx[2] / betas[0];
      betas[3] = -bx[3] / betas[0];
    } else {
      betas[0] = sqrt(bx[0]);
      betas[1] = bx[1] / betas[0];
      betas[2] = bx[2] / betas[0];
      betas[3] = bx[3] / betas[0];
    }
  } else {
    if (bx[0] < 0) {
      betas[0] = sqrt(-bx[0]);
      betas[1] = (bx[2] < 0) ? sqrt(-bx[2]) : 0.0;
    } else {
      betas[0] = sqrt(bx[0]);
      betas[1] = (bx[2] > 0) ? sqrt(bx[2]) : 0.0;
    }
    if (bx[1] < 0) betas[0] = -betas[0];
    betas[2] = N == 2 ? 0.0 : bx[3] / betas[0];
    betas[3] = 0.0;
  }
  epnp_gauss_newton(l_6x10, rho, betas);
  return epnp_compute_R_and_t(w, K, ut, betas, R, t);
}

VO_HDN double epnp_variant(int N, const EpnpWork& w, const Intrinsics& K, const double* ut, const double* l_6x10,
                           const double* rho, double* betas, double R[3][3], double t[3]) {
  const int ncol = epnp_variant_ncol(N);
  double at[30], sw[5], v[25], bx[5] = {0, 0, 0, 0, 0};
  epnp_variant_system(N, l_6x10, at);
  jacobi_svd_rt<6, 5>(at, sw, v, ncol);
  svd_backsubst_6xn(at, sw, v, rho, bx, ncol);
  return epnp_variant_back(N, w, K, ut, l_6x10, rho, bx, betas, R, t);
}

VO_HDN void epnp5_back(const EpnpWork& w, const Intrinsics& K, const double* ut, double Rout[9], double tout[3],
                       double* dbg = nullptr) {
  double l_6x10[60], rho[6];
  epnp_prepare(w, ut, l_6x10, rho);
  double Betas[4][4], rep_errors[4];
  double Rs[4][3][3], ts[4][3];
  for (int N = 1; N <= 3; N++) rep_errors[N] = epnp_variant(N, w, K, ut, l_6x10, rho, Betas[N], Rs[N], ts[N]);
  if (dbg) {
    for (int i = 0; i < 10; i++) dbg[i] = w.us[i];
    for (int i = 0; i < 12; i++) dbg[10 + i] = w.cws[i / 3][i % 3];
    for (int i = 0; i < 20; i++) dbg[22 + i] = w.alphas[i];
    for (int i = 0; i < 144; i++) dbg[186 + i] = ut[i];
    for (int i = 0; i < 60; i++) dbg[342 + i] = l_6x10[i];
    for (int i = 0; i < 6; i++) dbg[402 + i] = rho[i];
    for (int i = 0; i < 3; i++) dbg[408 + i] = rep_errors[1 + i];
    for (int i = 0; i < 4; i++) dbg[411 + i] = Betas[1][i];
    for (int i = 0; i < 4; i++) dbg[415 + i] = Betas[2][i];
  }
  int N = 1;
  if (rep_errors[2] < rep_errors[1]) N = 2;
  if (rep_errors[3] < rep_errors[N]) N = 3;
  for (int i = 0; i < 3; i++) {
    tout[i] = ts[N][i];
    for (int j = 0; j < 3; j++) Rout[i * 3 + j] = Rs[N][i][j];
  }
}

// Returns R (row-major) and t.  dbg (>= 420 doubles, optional): intermediates for parity debugging.
template <bool FMA_MTM>
VO_HDN void epnp5(const float* obj /*15*/, const float* img /*10*/, const Intrinsics& K, double Rout[9], double tout[3],
                  double* dbg = nullptr) {
  EpnpWork w;
  double mtm[144], d[12], ut[144];
  epnp5_front<FMA_MTM>(obj, img, K, w, mtm);
  if (dbg)
    for (int i = 0; i < 144; i++) dbg[42 + i] = mtm[i];
  svd_square_ut_only<12>(mtm, d, ut);
  if (dbg)
    for (int i = 0; i < 12; i++) dbg[330 + i] = d[i];
  epnp5_back(w, K, ut, Rout, tout, dbg);
}

// ---------------------------------------------------------------------------------------
// cv::triangulatePoints for one correspondence: DLT rows, Jacobi SVD of the 4x4, last row
// of V^T stored as float, then the reference's float dehomogenisation
// (src/triangulation.cpp:154-160).
VO_HD void triangulate_dlt(const double* P1, const double* P2, float x1, float y1, float x2, float y2, float out_xyz[3],
                           float* out_h4 = nullptr) {
  double A[16], w[4], vt[16];
  const double xs[2] = {(double)x1, (double)x2}, ys[2] = {(double)y1, (double)y2};
  const double* Ps[2] = {P1, P2};
  for (int j = 0; j < 2; j++)
    for (int k = 0; k < 4; k++) {
      A[(j * 2 + 0) * 4 + k] = xs[j] * Ps[j][8 + k] - Ps[j][k];
      A[(j * 2 + 1) * 4 + k] = ys[j] * Ps[j][8 + k] - Ps[j][4 + k];
    }
  svd_square<4>(A, w, nullptr, vt);
  float X = (float)vt[12], Y = (float)vt[13], Z = (float)vt[14], Wh = (float)vt[15];
  if (out_h4) {
    out_h4[0] = X; out_h4[1] = Y; out_h4[2] = Z; out_h4[3] = Wh;
  }
  out_xyz[0] = X / Wh;
  out_xyz[1] = Y / Wh;
  out_xyz[2] = Z / Wh;
}

}  // namespace vo
