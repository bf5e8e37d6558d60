// jacobi_warp.cuh -- the sweep phase of OpenCV's one-sided Jacobi SVD, executed by one warp
// with the SAME sequence of floating-point operations as the serial loop of cvmath.cuh.
//
// The serial algorithm visits the row pairs (0,1),(0,2),...,(0,N-1),(1,2),... ; a pair (i,j)
// reads and writes rows i and j only, so its result depends on the last earlier pair that
// touched row i and the last one that touched row j -- nothing else.  In the row-cyclic order
// that makes pair (i,j) of sweep s ready at wavefront step  s*N + i + j : all pairs on one
// anti-diagonal i + j = t are independent (disjoint rows), and sweep s+1 may start N steps after
// sweep s (its pair (i,j) needs the rows as left by (i,N-1) and (j,N-1) of sweep s, which ran at
// steps i+N-1 and j+N-1 < i+j+N).  One lane takes one pair; lanes [0,H) work on the older of the
// two sweeps in flight, lanes [H,2H) on the newer, H = N/2.  A 12x12 problem needs ~N steps per
// sweep instead of N(N-1)/2 = 66, and every step costs one rotation's latency (3 divides and 3
// square roots in FP64 -- the chain that makes the serial version latency-bound).
//
// OpenCV stops after the first sweep without a rotation.  Here the following sweep has already
// started speculatively; that is harmless: if a sweep changed nothing, every pair of the next
// sweep sees exactly the data its predecessor saw and skips as well.
#pragma once
#include <float.h>

#include "cvmath.cuh"

namespace vo {

// At: N rows of length M in shared memory (row-major, stride M), W: N squared row norms in
// shared memory (initialised by the caller exactly like the serial code: sequential sum of
// squares).  All 32 lanes must call.  On return the rows are orthogonal (not yet normalised /
// sorted: jacobi_svd<..., SKIP_SWEEPS=true> does that).
template <int M, int N, bool SPEC = false>
__device__ __forceinline__ void jacobi_sweeps_warp(double* At, double* W, int lane) {
  constexpr int H = N / 2;          // max independent pairs per anti-diagonal
  constexpr int P = N;              // steps between the starts of consecutive sweeps
  constexpr int D = 2 * N - 3;      // steps per sweep (anti-diagonals t = 1 .. 2N-3)
  constexpr int max_iter = M > 30 ? M : 30;
  static_assert(2 * H <= 32 && D > P && D < 2 * P, "schedule assumes two sweeps in flight");
  const double eps = DBL_EPSILON * 10;
  const bool is_new = lane >= H;
  const int idx = is_new ? lane - H : lane;
  bool chg_old = false, chg_new = false;   // warp-uniform: did the older / newer sweep rotate
  int s_new = 0;                           // index of the newer sweep
  for (int g = 1;; g++) {
    const int t_new = g - s_new * P;       // anti-diagonal of the newer sweep, 1..P
    const int sweep = is_new ? s_new : s_new - 1;
    const int t = is_new ? t_new : t_new + P;
    bool rotated = false;
    if (lane < 2 * H && sweep >= 0 && sweep < max_iter && t <= D) {
      const int i0 = t - (N - 1) > 0 ? t - (N - 1) : 0;
      const int i = i0 + idx, j = t - i;
      if (i < j) {
        double* Ai = At + i * M;
        double* Aj = At + j * M;
        double a = W[i], p = 0, b = W[j];
        double ri[M], rj[M];
#pragma unroll
        for (int k = 0; k < M; k++) {
          ri[k] = Ai[k];
          rj[k] = Aj[k];
        }
#pragma unroll
        for (int k = 0; k < M; k++) p += ri[k] * rj[k];
        // SPEC: the hypot of the rotation is started next to the skip test's square root (two independent chains)
        const double p2 = p * 2;
        const double beta = a - b;
        double gamma = 0;
        if (SPEC) gamma = cv_hypot(p2, beta);
        if (!(fabs(p) <= eps * sqrt(a * b))) {
          double c, s;
          p = p2;
          if (!SPEC) gamma = cv_hypot(p, beta);
          cv_jacobi_cs(p, beta, gamma, c, s);
          a = b = 0;
#pragma unroll
          for (int k = 0; k < M; k++) {
            const double t0 = c * ri[k] + s * rj[k];
            const double t1 = -s * ri[k] + c * rj[k];
            Ai[k] = t0;
            Aj[k] = t1;
            a += t0 * t0;
            b += t1 * t1;
          }
          W[i] = a;
          W[j] = b;
          rotated = true;
        }
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, rotated);
    __syncwarp();   // row / W updates visible before the next anti-diagonal
    chg_old |= (bal & ((1u << H) - 1u)) != 0;
    chg_new |= (bal >> H) != 0;
    // the older sweep finishes when its anti-diagonal reaches D
    if (s_new >= 1 && t_new + P == D) {
      if (!chg_old || s_new - 1 == max_iter - 1) break;
    }
    if (t_new == P) {   // next step opens a new sweep; the newer one becomes the older one
      chg_old = chg_new;
      chg_new = false;
      s_new++;
    }
  }
}

// The sweeps of up to four small runtime-size problems (M columns, n_g <= NMAX rows, V accumulated: jacobi_rt_rotate of
// cvmath.cuh) side by side: lane group g = lane / 8 works on problem g with the same wavefront schedule as above -- pair
// (i, j) of sweep s at step s * n + i + j, H = n / 2 lanes on the older and H on the newer sweep in flight (for n = 3 the
// pairs of a sweep are strictly sequential and one lane suffices).  EPnP's three beta initialisations are 6 x 4, 6 x 3
// and 6 x 5 systems: the 10 pairs per sweep of the largest take 5 steps instead of 10.  At / W / Vt: problem g at
// base + g * stride.  n_g = 0 marks an unused group.  All 32 lanes must call.
template <int M, int NMAX>
__device__ __forceinline__ void jacobi_sweeps_groups(double* At, int at_stride, double* W, int w_stride, double* Vt,
                                                     int v_stride, int n_g, int lane) {
  constexpr int max_iter = M > 30 ? M : 30;
  const int g = lane >> 3, idx = lane & 7;
  const int n = n_g;
  const int H = n >> 1, P = n, D = 2 * n - 3;
  double* At_g = At + g * at_stride;
  double* W_g = W + g * w_stride;
  double* Vt_g = Vt + g * v_stride;
  bool done = n < 2;
  bool chg_old = false, chg_new = false;    // group-uniform
  int s_new = 0;
  const unsigned half = (1u << H) - 1u;
  for (int step = 1;; step++) {
    const int t_new = step - s_new * P;      // anti-diagonal of the newer sweep, 1..P
    const bool is_new = idx >= H;
    const int sweep = is_new ? s_new : s_new - 1;
    const int t = is_new ? t_new : t_new + P;
    bool rotated = false;
    if (!done && idx < 2 * H && sweep >= 0 && sweep < max_iter && t <= D) {
      const int li = is_new ? idx - H : idx;
      const int i0 = t - (n - 1) > 0 ? t - (n - 1) : 0;
      const int i = i0 + li, j = t - i;
      if (i < j) rotated = jacobi_rt_rotate<M, NMAX>(At_g, W_g, Vt_g, i, j);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, rotated);
    __syncwarp();   // row / W / V updates visible before the next anti-diagonal
    if (!done) {
      const unsigned grp = (bal >> (8 * g)) & 0xffu;
      chg_old |= (grp & half) != 0;
      chg_new |= ((grp >> H) & half) != 0;
      if (D > P) {   // a sweep ends in the older role
        if (s_new >= 1 && t_new + P == D && (!chg_old || s_new - 1 == max_iter - 1)) done = true;
      } else {       // n = 3: it ends where the next one starts
        if (t_new == D && (!chg_new || s_new == max_iter - 1)) done = true;
      }
      if (t_new == P) {   // next step opens a new sweep; the newer one becomes the older one
        chg_old = chg_new;
        chg_new = false;
        s_new++;
      }
    }
    if (__all_sync(0xffffffffu, done)) break;
  }
}

}  // namespace vo
