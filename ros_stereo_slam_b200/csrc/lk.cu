// lk.cu -- K2: pyramidal Lucas-Kanade tracking, one warp per keypoint, all pyramid levels in one
// launch.  Replaces cv::calcOpticalFlowPyrLK as the reference calls it with all defaults
// (src/tracking.cpp:18 stereo L->R, :52 temporal): 21x21 window, 4 levels, 30 iterations / eps 0.01,
// minEigThreshold 1e-4.
//
// BIT-IDENTICAL to OpenCV 4.13 (SSE baseline build of video/lkpyramid.cpp), including the order in
// which OpenCV accumulates its window sums in float.  OpenCV walks a window row in steps of 8
// samples: the A sums (Ix*Ix, Ix*Iy, Iy*Iy) go through four float SIMD lanes (sample x -> lane x&3,
// product and sum each rounded), the b sums (diff*Ix, diff*Iy) through the same four lanes with the
// exact int32 pair sum of samples (x, x+4) converted to float first; the samples left over after the
// last full step of 8 (x = 16..20) go through ONE scalar float accumulator.  Per sum that is FIVE
// sequential float CHAINS over the whole window (row-major), combined at the end as
//     total = tail + ((c0 + c2) + (c1 + c3))          (every + rounded to float)
// oracle/lk.py restates this and is pinned bit for bit against cv2 (tests/test_oracle_lk.py).
//
// How the sequential chains are evaluated in parallel: every term is an INTEGER, so a chain is exact --
// and therefore order-independent -- as long as no partial sum reaches 2^24.  Each lane accumulates
// its samples in integers into per-chain registers (the chain of a sample is a compile-time property of
// its register), and next to them a bound on the sum of |term| per chain.  If the warp-wide bounds stay
// below 2^24 (98-99 % of the iterations on the synthetic sequences) the five exact chain totals go
// through the final float combination above and the result is OpenCV's, bit for bit.  Otherwise the warp
// takes the slow path: the float terms are written to shared memory and lanes 0..9 (0..4 per A sum) add
// them up one by one in OpenCV's order.
//
// Mapping: a window row is cut into two OCTS (samples 0..7, 8..15) and one TAIL (16..20).  Lane l owns
// oct l in slot A; in slot B lanes 0..9 own octs 32..41 and lanes 10..30 the tail of row l-10 (processed
// as an oct whose samples 5..7 have zero multiplicands).  In an oct, sample i belongs to chain i&3 and
// (i, i+4) is one of OpenCV's int32 pairs.  Bilinear samples are two DP2A each (signed 16-bit weight
// pair x unsigned 8-bit pixel pair), from a warp-private shared-memory tile.
//
// Memory path: the 28 x 32-byte J tile is staged with cp.async (row-coalesced: 8 lanes per row) around the
// window WITH A MARGIN of 3 px, so that the usual sub-pixel moves of an LK iteration never restage; the
// tile of a level is requested before the window extraction of that level starts and lands behind it.
// Levels live padded in HBM/L2 (common.cuh), so none of this carries bounds logic.
#include "common.cuh"

namespace vo {

namespace {

constexpr int WIN = LK_WIN;
constexpr int W_BITS = 14;
constexpr int TROWS = 28;                 // staged rows
constexpr int TS = 9;                     // tile row stride in words (8 used + 1: conflict-poor)
constexpr int TILE_WORDS = TROWS * TS;    // 252
constexpr int DROWS = WIN + 1;            // 22 derivative rows
constexpr int DS = 24;                    // derivative tile row stride (16-byte aligned rows)
constexpr int DTILE_WORDS = DROWS * DS + 8;   // 536 (the tail lanes read one word past the last row)
constexpr int NB_SIMD = 2 * WIN;          // 42 pair terms per SIMD chain of a b sum
constexpr int NA_SIMD = 4 * WIN;          // 84 terms per SIMD chain of an A sum
constexpr int N_TAIL = 5 * WIN;           // 105 terms of the scalar chain
constexpr int FB_WORDS = 2 * (4 * NB_SIMD + N_TAIL);   // 546: b1 | b2 terms
constexpr int FA_WORDS = 4 * NA_SIMD + N_TAIL;         // 441: one A sum at a time
constexpr int SCRATCH_WORDS = (TILE_WORDS + DTILE_WORDS) > FB_WORDS ? (TILE_WORDS + DTILE_WORDS) : FB_WORDS;  // 780
constexpr int WARP_WORDS = TILE_WORDS + SCRATCH_WORDS;  // J tile | {I tile + derivative tile} U {chain terms}
constexpr int WARP_WORDS_G3 = TILE_WORDS + 3 * LK_WIN * LK_WIN + 1;   // the replicated-gray variant keeps whole windows of terms
                                          // for its fall-backs: three A sums (3 x 441 floats) / two b sums (2 x 441 int32)
#ifndef LK_WARPS_N
#define LK_WARPS_N 4
#endif
constexpr int LK_WARPS = LK_WARPS_N;
constexpr int MARGIN = 3;
constexpr int SAFE_LIMIT = 1 << 24;
constexpr int BOUND_CLAMP = 1 << 26;    // per-lane clamp of a bound before the warp sum (32 x 2^26 fits 32 bits)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void lk_weights(float a, float b, unsigned& wt, unsigned& wb, int& iw00, int& iw01, int& iw10,
                                           int& iw11) {
  const float s = (float)(1 << W_BITS);
  const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
  iw00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), s));
  iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), s));
  iw10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), s));
  iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
  wt = ((unsigned)iw00 & 0xffffu) | ((unsigned)iw01 << 16);   // weights are in [-1, 2^14]: two s16 per register
  wb = ((unsigned)iw10 & 0xffffu) | ((unsigned)iw11 << 16);
}

// d = c + a.s16[0]*b.u8[0] + a.s16[1]*b.u8[1]   (.hi: bytes 2,3 of b)
__device__ __forceinline__ int dp2a_lo(unsigned w, unsigned px, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi(unsigned w, unsigned px, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c));
  return d;
}

__device__ __forceinline__ void cp_async4(unsigned* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Request the 28 x 8-word tile whose first word is at `src` (4-byte aligned): lane -> (row 4i + lane/8, word lane%8).
__device__ __forceinline__ void stage_tile(unsigned* tile_lane, const uint8_t* src, int pitch, int lane) {
  const unsigned* w = reinterpret_cast<const unsigned*>(src) + (lane >> 3) * (pitch >> 2) + (lane & 7);
#pragma unroll
  for (int i = 0; i < TROWS / 4; i++) {
    cp_async4(tile_lane + i * 4 * TS, w);
    w += pitch;   // 4 rows, in words
  }
}

// the 8 bilinear samples of the oct whose first source byte is `sh`/8 bytes into the tile word p[0] (and the row below)
__device__ __forceinline__ void oct_sample(const unsigned* p, unsigned sh, unsigned wt, unsigned wb, int out[8]) {
  const unsigned a0 = p[0], a1 = p[1], a2 = p[2];
  const unsigned c0 = p[TS], c1 = p[TS + 1], c2 = p[TS + 2];
  const unsigned t0 = __funnelshift_r(a0, a1, sh), t1 = __funnelshift_r(a1, a2, sh), t2 = a2 >> sh;
  const unsigned b0 = __funnelshift_r(c0, c1, sh), b1 = __funnelshift_r(c1, c2, sh), b2 = c2 >> sh;
  // pixel pairs (x, x+1): even x from t0/t1 via dp2a.lo/.hi, odd x from the registers shifted by one byte
  const unsigned tu = __funnelshift_r(t0, t1, 8), tv = __funnelshift_r(t1, t2, 8);
  const unsigned bu = __funnelshift_r(b0, b1, 8), bv = __funnelshift_r(b1, b2, 8);
  const int rc = 1 << (W_BITS - 5 - 1);
  out[0] = dp2a_lo(wb, b0, dp2a_lo(wt, t0, rc)) >> (W_BITS - 5);
  out[1] = dp2a_lo(wb, bu, dp2a_lo(wt, tu, rc)) >> (W_BITS - 5);
  out[2] = dp2a_hi(wb, b0, dp2a_hi(wt, t0, rc)) >> (W_BITS - 5);
  out[3] = dp2a_hi(wb, bu, dp2a_hi(wt, tu, rc)) >> (W_BITS - 5);
  out[4] = dp2a_lo(wb, b1, dp2a_lo(wt, t1, rc)) >> (W_BITS - 5);
  out[5] = dp2a_lo(wb, bv, dp2a_lo(wt, tv, rc)) >> (W_BITS - 5);
  out[6] = dp2a_hi(wb, b1, dp2a_hi(wt, t1, rc)) >> (W_BITS - 5);
  out[7] = dp2a_hi(wb, bv, dp2a_hi(wt, tv, rc)) >> (W_BITS - 5);
}

// bilinear (Ix, Iy) of the 8 samples of an oct from the staged short2 derivative tile (row pointer d, 9 columns)
__device__ __forceinline__ void oct_deriv(const unsigned* d, int iw00, int iw01, int iw10, int iw11, int ix[8], int iy[8]) {
  unsigned top_[9], bot_[9];
  {
    const uint4 q0 = *reinterpret_cast<const uint4*>(d), q1 = *reinterpret_cast<const uint4*>(d + 4);
    const uint4 r0 = *reinterpret_cast<const uint4*>(d + DS), r1 = *reinterpret_cast<const uint4*>(d + DS + 4);
    top_[0] = q0.x; top_[1] = q0.y; top_[2] = q0.z; top_[3] = q0.w; top_[4] = q1.x; top_[5] = q1.y; top_[6] = q1.z; top_[7] = q1.w;
    bot_[0] = r0.x; bot_[1] = r0.y; bot_[2] = r0.z; bot_[3] = r0.w; bot_[4] = r1.x; bot_[5] = r1.y; bot_[6] = r1.z; bot_[7] = r1.w;
    top_[8] = d[8];
    bot_[8] = d[DS + 8];
  }
#pragma unroll
  for (int x = 0; x < 8; x++) {
    // short2 packed in a word: .x = low half (dx), .y = high half (dy)
    ix[x] = ((int)(short)(top_[x] & 0xffff) * iw00 + (int)(short)(top_[x + 1] & 0xffff) * iw01 +
             (int)(short)(bot_[x] & 0xffff) * iw10 + (int)(short)(bot_[x + 1] & 0xffff) * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
    iy[x] = (((int)top_[x] >> 16) * iw00 + ((int)top_[x + 1] >> 16) * iw01 + ((int)bot_[x] >> 16) * iw10 +
             ((int)bot_[x + 1] >> 16) * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
  }
}

// OpenCV's final combination of the five chains of one sum (every + rounded to float)
__device__ __forceinline__ float chain_combine(float c0, float c1, float c2, float c3, float tail) {
  return __fadd_rn(tail, __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3)));
}

// per-lane chain totals of one sum: slot A (always an oct) + slot B (an oct for !isq, a tail for isq)
__device__ __forceinline__ void lane_chains(const int a[4], const int b[4], int notq, int isq, int c[5]) {
#pragma unroll
  for (int k = 0; k < 4; k++) c[k] = a[k] + notq * b[k];
  c[4] = isq * ((b[0] + b[1]) + (b[2] + b[3]));
}

// per-lane chain totals of one sum, 3-channel image with identical planes (see lk_kernel<.., true>): a[j] = slot A's
// column class j; b[i] = slot B's sample i (an oct's column class i & 3, or a tail's column 16 + i)
template <bool G3, int NB>
__device__ __forceinline__ void lane_chains_t(const int a[4], const int (&b)[NB], int notq, int isq, int c[5]) {
  if constexpr (!G3) {
    lane_chains(a, b, notq, isq, c);
  } else {
    int Q[4], T[5];
#pragma unroll
    for (int j = 0; j < 4; j++) Q[j] = a[j] + notq * (b[j] + b[j + 4]);
#pragma unroll
    for (int i = 0; i < 5; i++) T[i] = isq * b[i];
    c[0] = (Q[0] + Q[1]) + (Q[2] + T[0]) + T[1];
    c[1] = (Q[0] + Q[1]) + (Q[3] + T[0]) + T[1];
    c[2] = (Q[0] + Q[2]) + (Q[3] + T[0]) + T[2];
    c[3] = (Q[1] + Q[2]) + (Q[3] + T[1]) + T[2];
    c[4] = T[2] + 3 * (T[3] + T[4]);
  }
}

// G3 fall-backs: the window's terms in the plain [row][col] layout; lane k < 4 adds the SIMD chain k (interleaved
// samples x = k, k+4, ... < 56 of every row, column x / 3), lane 4 the scalar chain (x = 56..62)
__device__ __forceinline__ float run_chain_g3_a(const float* f, int lane) {
  float acc = 0.f;
  if (lane < 4) {
    for (int r = 0; r < WIN; r++)
      for (int x = lane; x < 56; x += 4) acc = __fadd_rn(acc, f[r * WIN + x / 3]);
  } else if (lane == 4) {
    for (int r = 0; r < WIN; r++)
      for (int x = 56; x < 63; x++) acc = __fadd_rn(acc, f[r * WIN + x / 3]);
  }
  return acc;
}
// A sums: the three sums side by side (windows at f, f + 441, f + 882), lanes 0..4 / 5..9 / 10..14
__device__ __forceinline__ float run_chain_g3_a3(const float* f, int lane) {
  float acc = 0.f;
  if (lane < 15) {
    const int s = lane / 5, k = lane - 5 * s;
    const float* q = f + s * WIN * WIN;
    if (k < 4) {
      for (int r = 0; r < WIN; r++)
        for (int x = k; x < 56; x += 4) acc = __fadd_rn(acc, q[r * WIN + x / 3]);
    } else {
      for (int r = 0; r < WIN; r++)
        for (int x = 56; x < 63; x++) acc = __fadd_rn(acc, q[r * WIN + x / 3]);
    }
  }
  return acc;
}

// Exact test of the five chains of ONE b sum (window of int32 products at P, plain layout): lane r < 21 holds row r,
// builds the 7 elements of every chain from registers (SIMD chain k: OpenCV's pairs x = 8t + k and x + 4; scalar chain:
// x = 56 + t; column = x / 3), a warp scan gives the exact partial sum in front of every row.  tot[k] = chain totals;
// returns the largest |partial sum| or |element| seen (clamped), i.e. < 2^24 iff all five float chains are exact.
__device__ __forceinline__ int g3_chains_exact(const int* P, int lane, int tot[5]) {
  int v[WIN];
  const int* row = P + (lane < WIN ? lane : 0) * WIN;
#pragma unroll
  for (int cidx = 0; cidx < WIN; cidx++) v[cidx] = lane < WIN ? row[cidx] : 0;
  int worst = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    int e[7];
    int run = 0;
#pragma unroll
    for (int t = 0; t < 7; t++) {
      e[t] = k < 4 ? v[(8 * t + k) / 3] + v[(8 * t + k + 4) / 3] : v[(56 + t) / 3];
      run += e[t];
    }
    int incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += u;
    }
    int p = incl - run;
#pragma unroll
    for (int t = 0; t < 7; t++) {
      p += e[t];
      worst = max(worst, max(abs(p), abs(e[t])));
    }
    tot[k] = __shfl_sync(FULL, incl, 31);
  }
  return __reduce_max_sync(FULL, min(worst, SAFE_LIMIT));
}

// b sums: lanes 0..4 = chains of b1 (products at p), lanes 5..9 = chains of b2 (products at p + 441); OpenCV's int32
// pairs (x, x + 4) inside every step of 8
__device__ __forceinline__ float run_chain_g3_b(const int* p, int lane) {
  float acc = 0.f;
  if (lane < 10) {
    const int* q = p + (lane >= 5 ? WIN * WIN : 0);
    const int k = lane >= 5 ? lane - 5 : lane;
    if (k < 4) {
      for (int r = 0; r < WIN; r++)
        for (int x = k; x < 56; x += 8) acc = __fadd_rn(acc, __int2float_rn(q[r * WIN + x / 3] + q[r * WIN + (x + 4) / 3]));
    } else {
      for (int r = 0; r < WIN; r++)
        for (int x = 56; x < 63; x++) acc = __fadd_rn(acc, __int2float_rn(q[r * WIN + x / 3]));
    }
  }
  return acc;
}

// lanes 0 .. nsum*5-1 add up their chain of float terms one by one (OpenCV's order); terms of sum s start at
// f + s*stride: four SIMD chains of n_simd terms, then the tail chain of N_TAIL terms.  Returns the five
// chain values of sum `s` broadcast to every lane.
__device__ __forceinline__ float run_chain(const float* f, int lane, int nsum, int n_simd) {
  float acc = 0.f;
  if (lane < nsum * 5) {
    const int s = lane / 5, k = lane - s * 5;
    const float* t = f + s * (4 * n_simd + N_TAIL) + (k < 4 ? k * n_simd : 4 * n_simd);
    const int n = k < 4 ? n_simd : N_TAIL;
    for (int j = 0; j < n; j++) acc = __fadd_rn(acc, t[j]);
  }
  return acc;
}

}  // namespace

// G3 = true: 3-channel images whose planes are identical (PyrView::mono of both images) -- see the note above
// lk_kernel_c3.  Every interleaved sample x = 3*col + ch then equals the 1-channel sample of its column, so ONE warp
// on plane 0 knows all 63 x 21 terms: column class j = col & 3 (cols 0..15) feeds the SIMD chains (3j + ch) & 3,
// columns 16 / 17 / 18 feed chains {0,1,2} / {3,0,1} / {2,3,scalar}, columns 19 and 20 feed the scalar chain three
// times each; the sequential fall-back walks the interleaved order over the 21-column term buffer.
template <int MINB, bool G3>
__global__ void __launch_bounds__(LK_WARPS * 32, MINB * 4 / LK_WARPS)
lk_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts, int n, float2* __restrict__ next_pts,
          uint8_t* __restrict__ status, float* __restrict__ err, int max_iters, double eps_sq, float eps_lo, float eps_hi,
          float min_eig_thr, unsigned long long* __restrict__ work, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  if (G3 && !(*prev.mono && *next.mono)) return;     // not a replicated-gray pair: lk_kernel_c3 does the launch
  constexpr int WW = G3 ? WARP_WORDS_G3 : WARP_WORDS;
  constexpr int CN = G3 ? 3 : 1;
  constexpr int NB = G3 ? 8 : 4;              // slot-B accumulators: per sample (G3: a tail's columns go to different chains)
  __shared__ __align__(16) unsigned smem[LK_WARPS * WW];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  unsigned* jtile = smem + (threadIdx.x >> 5) * WW;
  unsigned* itile = jtile + TILE_WORDS;        // scratch: I tile | derivative tile, later the chain terms
  unsigned* dtile = itile + TILE_WORDS;
  float* fterms = reinterpret_cast<float*>(itile);
  const int st_lane = (lane >> 3) * TS + (lane & 7);
  const float2 pt = prev_pts[warp];
  const float half_win = (WIN - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);

  // this lane's two units (fixed for the whole kernel).  The per-lane constants are packed into one register
  // behind an opaque move: under register pressure ptxas otherwise re-derives them from %tid in every iteration.
  unsigned cfg;
  {
    const int rA = lane >> 1, cA = (lane & 1) * 8;
    const bool q = lane >= 10, hb = lane < 31;
    const int rB = q ? (hb ? lane - 10 : 0) : 16 + (lane >> 1);
    const int cB = q ? 16 : (lane & 1) * 8;
    const unsigned v = (unsigned)(rA * TS + (cA >> 2)) | ((unsigned)(rB * TS + (cB >> 2)) << 8) | ((unsigned)rA << 16) |
                       ((unsigned)rB << 21) | ((unsigned)(cA >> 3) << 26) | ((unsigned)(cB >> 3) << 27) | (q ? 1u << 29 : 0u) |
                       (hb ? 1u << 30 : 0u);
    asm volatile("mov.b32 %0, %1;" : "=r"(cfg) : "r"(v));
  }
#define offA ((int)(cfg & 0xffu))            /* word offsets of the two units in a tile */
#define offB ((int)((cfg >> 8) & 0xffu))
#define rowA ((int)((cfg >> 16) & 31u))
#define rowB ((int)((cfg >> 21) & 31u))
#define colA ((int)((cfg >> 26) & 1u) * 8)
#define colB ((int)((cfg >> 27) & 3u) * 8)
#define isq_b ((cfg & (1u << 29)) != 0)     /* slot B is a tail (or, lane 31, nothing) */
#define hasB ((cfg & (1u << 30)) != 0)
#define isq ((int)((cfg >> 29) & 1u))
#define notq (1 - isq)

  float outx = 0.f, outy = 0.f;  // nextPts[ptidx] as OpenCV keeps it between levels
  bool st = true;
  float errv = 0.f;
  unsigned int n_levels_done = 0, n_iters_done = 0, n_slow_a = 0, n_slow_b = 0;

  int Iw[16], Ix[16], Iy[16];
  int mpA[4], mB[8];

  const int top = prev.nlevels - 1;
  for (int level = top; level >= 0; level--) {
    const PyrLevelView I = prev.lv[level];
    const PyrLevelView J = next.lv[level];
    const int pitch = I.pitch;
    const float scale = __int_as_float((127 - level) << 23);   // 2^-level
    float px = __fmul_rn(pt.x, scale), py = __fmul_rn(pt.y, scale);
    float nx, ny;
    if (level == top) {
      nx = px;
      ny = py;
    } else {
      nx = __fmul_rn(outx, 2.f);
      ny = __fmul_rn(outy, 2.f);
    }
    outx = nx;
    outy = ny;

    px = __fsub_rn(px, half_win);
    py = __fsub_rn(py, half_win);
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -WIN || ipx >= I.w || ipy < -WIN || ipy >= I.h) {
      if (level == 0) {
        st = false;
        errv = 0.f;
      }
      continue;
    }
    int iw00, iw01, iw10, iw11;
    unsigned wt, wb;
    lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), wt, wb, iw00, iw01, iw10, iw11);

    // ---- request the I tile, the derivative patch and (with a margin) the J tile of the start position
    nx = __fsub_rn(nx, half_win);
    ny = __fsub_rn(ny, half_win);
    int tX0 = -(1 << 20), tY0 = -(1 << 20);     // origin of the staged J tile in padded coordinates (none yet)
    __syncwarp();                                // everyone is done with the tiles of the previous level
    const int iX = ipx + PAD_L, iY = ipy + PAD_Y;
    {
      stage_tile(itile + st_lane, I.img + (size_t)iY * pitch + (iX & ~3), pitch, lane);
      // derivative patch: 22 rows x 24 short2 (22 used), 8 lanes x 3 words per row, 4 rows per step
      {
        const unsigned* dsrc = reinterpret_cast<const unsigned*>(I.deriv + (size_t)iY * pitch + iX) + (lane >> 3) * pitch + (lane & 7);
        unsigned* ddst = dtile + (lane >> 3) * DS + (lane & 7);
#pragma unroll
        for (int i = 0; i < 6; i++) {
          if (i < 5 || (lane >> 3) < DROWS - 20) {
            cp_async4(ddst, dsrc);
            cp_async4(ddst + 8, dsrc + 8);
            cp_async4(ddst + 16, dsrc + 16);
          }
          dsrc += 4 * pitch;
          ddst += 4 * DS;
        }
      }
      cp_async_commit();
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (!(inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h)) {
        tX0 = (inx + PAD_L - MARGIN) & ~3;
        tY0 = iny + PAD_Y - MARGIN;
        stage_tile(jtile + st_lane, J.img + (size_t)tY0 * pitch + tX0, pitch, lane);
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
    }

    // ---- window extraction from the previous image + its Scharr derivative
    int cA11[5], cA12[5], cA22[5], cC1[5], cC2[5];
    {
      const unsigned shI = (unsigned)(iX & 3) * 8;
      oct_sample(itile + offA, shI, wt, wb, Iw);
      oct_sample(itile + offB, shI, wt, wb, Iw + 8);
      oct_deriv(dtile + rowA * DS + colA, iw00, iw01, iw10, iw11, Ix, Iy);
      oct_deriv(dtile + rowB * DS + colB, iw00, iw01, iw10, iw11, Ix + 8, Iy + 8);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const bool valid = hasB && (!isq_b || i < 5);
        if (!valid) { Iw[8 + i] = 0; Ix[8 + i] = 0; Iy[8 + i] = 0; }
      }
      int a11[4] = {0, 0, 0, 0}, a12[4] = {0, 0, 0, 0}, a22[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
      int b11[NB] = {0}, b12[NB] = {0}, b22[NB] = {0}, d1[NB] = {0}, d2[NB] = {0};
      int mA[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int k = i & 3;
        a11[k] += Ix[i] * Ix[i];
        a12[k] += Ix[i] * Iy[i];
        a22[k] += Iy[i] * Iy[i];
        c1[k] += Iw[i] * Ix[i];
        c2[k] += Iw[i] * Iy[i];
        mA[i] = max(abs(Ix[i]), abs(Iy[i]));
        const int kb = G3 ? i : k;
        b11[kb] += Ix[8 + i] * Ix[8 + i];
        b12[kb] += Ix[8 + i] * Iy[8 + i];
        b22[kb] += Iy[8 + i] * Iy[8 + i];
        d1[kb] += Iw[8 + i] * Ix[8 + i];
        d2[kb] += Iw[8 + i] * Iy[8 + i];
        mB[i] = max(abs(Ix[8 + i]), abs(Iy[8 + i]));
      }
#pragma unroll
      for (int k = 0; k < 4; k++) mpA[k] = max(mA[k], mA[k + 4]);
      lane_chains_t<G3>(a11, b11, notq, isq, cA11);
      lane_chains_t<G3>(a12, b12, notq, isq, cA12);
      lane_chains_t<G3>(a22, b22, notq, isq, cA22);
      lane_chains_t<G3>(c1, d1, notq, isq, cC1);
      lane_chains_t<G3>(c2, d2, notq, isq, cC2);
    }
    bool safeA = true;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      if (G3) {   // up to 294 squares per chain: the lane parts are clamped so that the warp sum cannot wrap (a clamped
                  // value only ever feeds the "unsafe" decision)
        cA11[k] = min(cA11[k], SAFE_LIMIT);
        cA22[k] = min(cA22[k], SAFE_LIMIT);
      }
      cA11[k] = __reduce_add_sync(FULL, cA11[k]);
      cA12[k] = __reduce_add_sync(FULL, cA12[k]);
      cA22[k] = __reduce_add_sync(FULL, cA22[k]);
      cC1[k] = __reduce_add_sync(FULL, cC1[k]);     // modulo 2^32: only differences with the per-iteration sums are used
      cC2[k] = __reduce_add_sync(FULL, cC2[k]);
      // the terms of the A11 / A22 chains are squares, so the chain totals (<= 105 * 4080^2 < 2^31) bound every
      // partial sum; |Ix*Iy| <= max(Ix^2, Iy^2) covers the A12 chain
      safeA = safeA && cA11[k] < SAFE_LIMIT && cA22[k] < SAFE_LIMIT;
    }
    float A11, A12, A22;
    if (safeA) {
      A11 = chain_combine((float)cA11[0], (float)cA11[1], (float)cA11[2], (float)cA11[3], (float)cA11[4]);
      A12 = chain_combine((float)cA12[0], (float)cA12[1], (float)cA12[2], (float)cA12[3], (float)cA12[4]);
      A22 = chain_combine((float)cA22[0], (float)cA22[1], (float)cA22[2], (float)cA22[3], (float)cA22[4]);
    } else {
      // slow path: the float terms in OpenCV's order (the I / derivative tiles are consumed)
      n_slow_a++;
      float res[3];
      if (G3) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) {
          fterms[rowA * WIN + colA + i] = __int2float_rn(Ix[i] * Ix[i]);
          fterms[WIN * WIN + rowA * WIN + colA + i] = __int2float_rn(Ix[i] * Iy[i]);
          fterms[2 * WIN * WIN + rowA * WIN + colA + i] = __int2float_rn(Iy[i] * Iy[i]);
          if (hasB && (!isq_b || i < 5)) {
            fterms[rowB * WIN + colB + i] = __int2float_rn(Ix[8 + i] * Ix[8 + i]);
            fterms[WIN * WIN + rowB * WIN + colB + i] = __int2float_rn(Ix[8 + i] * Iy[8 + i]);
            fterms[2 * WIN * WIN + rowB * WIN + colB + i] = __int2float_rn(Iy[8 + i] * Iy[8 + i]);
          }
        }
        __syncwarp();
        const float acc = run_chain_g3_a3(fterms, lane);
#pragma unroll
        for (int s = 0; s < 3; s++)
          res[s] = chain_combine(__shfl_sync(FULL, acc, 5 * s), __shfl_sync(FULL, acc, 5 * s + 1), __shfl_sync(FULL, acc, 5 * s + 2),
                                 __shfl_sync(FULL, acc, 5 * s + 3), __shfl_sync(FULL, acc, 5 * s + 4));
      } else {
  #pragma unroll
        for (int s = 0; s < 3; s++) {
          __syncwarp();
  #pragma unroll
          for (int i = 0; i < 8; i++) {
            const int pa = s == 0 ? Ix[i] * Ix[i] : (s == 1 ? Ix[i] * Iy[i] : Iy[i] * Iy[i]);
            const int pb = s == 0 ? Ix[8 + i] * Ix[8 + i] : (s == 1 ? Ix[8 + i] * Iy[8 + i] : Iy[8 + i] * Iy[8 + i]);
            if (G3) {   // plain window layout; the chain lanes walk the interleaved order over it
              fterms[rowA * WIN + colA + i] = __int2float_rn(pa);
              if (hasB && (!isq_b || i < 5)) fterms[rowB * WIN + colB + i] = __int2float_rn(pb);
            } else {
              fterms[(i & 3) * NA_SIMD + 4 * rowA + (colA >> 2) + (i >> 2)] = __int2float_rn(pa);
              if (!isq_b) fterms[(i & 3) * NA_SIMD + 4 * rowB + (colB >> 2) + (i >> 2)] = __int2float_rn(pb);
              else if (hasB && i < 5) fterms[4 * NA_SIMD + 5 * rowB + i] = __int2float_rn(pb);
            }
          }
          __syncwarp();
          const float acc = run_chain(fterms, lane, 1, NA_SIMD);
          res[s] = chain_combine(__shfl_sync(FULL, acc, 0), __shfl_sync(FULL, acc, 1), __shfl_sync(FULL, acc, 2),
                                 __shfl_sync(FULL, acc, 3), __shfl_sync(FULL, acc, 4));
        }
      }
      A11 = res[0];
      A12 = res[1];
      A22 = res[2];
    }
    const int C1tot = (int)((unsigned)cC1[0] + (unsigned)cC1[1] + (unsigned)cC1[2] + (unsigned)cC1[3] + (unsigned)cC1[4]);
    const int C2tot = (int)((unsigned)cC2[0] + (unsigned)cC2[1] + (unsigned)cC2[2] + (unsigned)cC2[3] + (unsigned)cC2[4]);
    A11 = __fmul_rn(A11, FLT_SCALE);
    A12 = __fmul_rn(A12, FLT_SCALE);
    A22 = __fmul_rn(A22, FLT_SCALE);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float q = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(q)), (float)(2 * WIN * WIN));
    n_levels_done++;
    cp_async_wait<0>();
    __syncwarp();
    if (min_eig < min_eig_thr || D < 1.1920929e-07f) {
      if (level == 0) st = false;
      continue;
    }
    D = __fdiv_rn(1.f, D);

    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_iters; j++) {
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
        if (level == 0) st = false;
        break;
      }
      lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), wt, wb, iw00, iw01, iw10, iw11);
      int bx = inx + PAD_L - tX0, by = iny + PAD_Y - tY0;
      if ((unsigned)bx > 10u || (unsigned)by > 6u) {   // the window left the staged tile
        __syncwarp();
        tX0 = (inx + PAD_L - MARGIN) & ~3;
        tY0 = iny + PAD_Y - MARGIN;
        stage_tile(jtile + st_lane, J.img + (size_t)tY0 * pitch + tX0, pitch, lane);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        bx = inx + PAD_L - tX0;
        by = MARGIN;
      }
      const unsigned* jp = jtile + by * TS + (bx >> 2);
      const unsigned shJ = (unsigned)(bx & 3) * 8;
      int c1[5], c2[5], cu[5];
      int t1, t2, tu;
      bool all_exact = true;
      {
        int a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0}, ua[4];
        int b1[NB] = {0}, b2[NB] = {0}, ub[NB] = {0};
        int jv[8];
        oct_sample(jp + offA, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          a1[i & 3] += jv[i] * Ix[i];
          a2[i & 3] += jv[i] * Iy[i];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) ua[k] = (int)__sad(jv[k + 4], Iw[k + 4], __sad(jv[k], Iw[k], 0u)) * mpA[k];
        oct_sample(jp + offB, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          b1[G3 ? i : (i & 3)] += jv[i] * Ix[8 + i];
          b2[G3 ? i : (i & 3)] += jv[i] * Iy[8 + i];
          ub[G3 ? i : (i & 3)] += (int)__sad(jv[i], Iw[8 + i], 0u) * mB[i];
        }
        t1 = ((a1[0] + a1[1]) + (a1[2] + a1[3])) + ((b1[0] + b1[1]) + (b1[2] + b1[3]));
        t2 = ((a2[0] + a2[1]) + (a2[2] + a2[3])) + ((b2[0] + b2[1]) + (b2[2] + b2[3]));
        tu = ((ua[0] + ua[1]) + (ua[2] + ua[3])) + ((ub[0] + ub[1]) + (ub[2] + ub[3]));   // < 2^30
        if (G3) {       // every sample stands for its three channels (tu <= 16 * 8160 * 4080 * 3 < 2^31)
          t1 = 3 * (t1 + ((b1[4] + b1[5]) + (b1[6] + b1[7])));
          t2 = 3 * (t2 + ((b2[4] + b2[5]) + (b2[6] + b2[7])));
          tu = 3 * (tu + ((ub[4] + ub[5]) + (ub[6] + ub[7])));
        }
        // Exactness bound.  With U >= sum |term| and T = sum term, the positive terms add up to at most (U + T) / 2 and
        // the negative ones to at most (U - T) / 2, so EVERY partial sum of the terms, in any order and grouping (the
        // chains, OpenCV's int32 pairs, the final combination), lies within +-(U + |T|) / 2: all float additions are
        // exact iff U + |T| < 2^25.  (The lane parts are clamped so that the warp sum cannot wrap; a clamped lane
        // fails the test by itself.  T is computed modulo 2^32 and is exact whenever the test can pass.)
        const unsigned U = __reduce_add_sync(FULL, (unsigned)min(tu, BOUND_CLAMP));
        t1 = (int)((unsigned)__reduce_add_sync(FULL, t1) - (unsigned)C1tot);
        t2 = (int)((unsigned)__reduce_add_sync(FULL, t2) - (unsigned)C2tot);
        if (U + (unsigned)max(abs(t1), abs(t2)) >= 2u * SAFE_LIMIT || U >= 2u * SAFE_LIMIT) {
          lane_chains_t<G3>(a1, b1, notq, isq, c1);
          lane_chains_t<G3>(a2, b2, notq, isq, c2);
          lane_chains_t<G3>(ua, ub, notq, isq, cu);
          all_exact = false;
        }
      }
      n_iters_done++;
      float b1f, b2f;
      bool safe = true;
      // tier 1 (all_exact): every chain and every step of the final combination is exact, the result is float(total)
      if (all_exact) {
        b1f = (float)t1;
        b2f = (float)t2;
      } else {
#pragma unroll
        for (int k = 0; k < 5; k++) {
          c1[k] = (int)((unsigned)__reduce_add_sync(FULL, c1[k]) - (unsigned)cC1[k]);   // sum(diff*Ix) = sum(J*Ix) - sum(I*Ix), modulo 2^32
          c2[k] = (int)((unsigned)__reduce_add_sync(FULL, c2[k]) - (unsigned)cC2[k]);
          const unsigned Uk = __reduce_add_sync(FULL, (unsigned)min(cu[k], BOUND_CLAMP));
          safe = safe && Uk < 2u * SAFE_LIMIT && Uk + (unsigned)max(abs(c1[k]), abs(c2[k])) < 2u * SAFE_LIMIT;
        }
      }
      if (all_exact) {
      } else if (safe) {
        b1f = chain_combine((float)c1[0], (float)c1[1], (float)c1[2], (float)c1[3], (float)c1[4]);
        b2f = chain_combine((float)c2[0], (float)c2[1], (float)c2[2], (float)c2[3], (float)c2[4]);
      } else {
        // slow path: float(pair sums) / float(tail products) to shared memory, lanes 0..9 add them in OpenCV's order
        n_slow_b++;
        int jv[8];
        __syncwarp();
        if (G3) {
          // the int32 products of the window in its plain layout (b1 | b2); the chain lanes form OpenCV's pairs
          // (x, x + 4) of the interleaved order themselves
          int* pt = reinterpret_cast<int*>(fterms);
          oct_sample(jp + offA, shJ, wt, wb, jv);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int d = jv[i] - Iw[i];
            pt[rowA * WIN + colA + i] = d * Ix[i];
            pt[WIN * WIN + rowA * WIN + colA + i] = d * Iy[i];
          }
          oct_sample(jp + offB, shJ, wt, wb, jv);
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (hasB && (!isq_b || i < 5)) {
              const int d = jv[i] - Iw[8 + i];
              pt[rowB * WIN + colB + i] = d * Ix[8 + i];
              pt[WIN * WIN + rowB * WIN + colB + i] = d * Iy[8 + i];
            }
          __syncwarp();
          // the bound on sum |term| failed: test the partial sums themselves (exactly) before falling back to the
          // sequential chains
          int e1[5], e2[5];
          const int w1 = g3_chains_exact(pt, lane, e1);
          const int w2 = g3_chains_exact(pt + WIN * WIN, lane, e2);
          if (max(w1, w2) < SAFE_LIMIT) {
            b1f = chain_combine((float)e1[0], (float)e1[1], (float)e1[2], (float)e1[3], (float)e1[4]);
            b2f = chain_combine((float)e2[0], (float)e2[1], (float)e2[2], (float)e2[3], (float)e2[4]);
          } else {
            const float acc = run_chain_g3_b(pt, lane);
            b1f = chain_combine(__shfl_sync(FULL, acc, 0), __shfl_sync(FULL, acc, 1), __shfl_sync(FULL, acc, 2),
                                __shfl_sync(FULL, acc, 3), __shfl_sync(FULL, acc, 4));
            b2f = chain_combine(__shfl_sync(FULL, acc, 5), __shfl_sync(FULL, acc, 6), __shfl_sync(FULL, acc, 7),
                                __shfl_sync(FULL, acc, 8), __shfl_sync(FULL, acc, 9));
          }
        } else {
          oct_sample(jp + offA, shJ, wt, wb, jv);
  #pragma unroll
          for (int k = 0; k < 4; k++) {
            const int dA = jv[k] - Iw[k], dB = jv[k + 4] - Iw[k + 4];
            fterms[k * NB_SIMD + 2 * rowA + (colA >> 3)] = __int2float_rn(dA * Ix[k] + dB * Ix[k + 4]);
            fterms[FB_WORDS / 2 + k * NB_SIMD + 2 * rowA + (colA >> 3)] = __int2float_rn(dA * Iy[k] + dB * Iy[k + 4]);
          }
          oct_sample(jp + offB, shJ, wt, wb, jv);
          if (!isq_b) {
  #pragma unroll
            for (int k = 0; k < 4; k++) {
              const int dA = jv[k] - Iw[8 + k], dB = jv[k + 4] - Iw[12 + k];
              fterms[k * NB_SIMD + 2 * rowB + (colB >> 3)] = __int2float_rn(dA * Ix[8 + k] + dB * Ix[12 + k]);
              fterms[FB_WORDS / 2 + k * NB_SIMD + 2 * rowB + (colB >> 3)] = __int2float_rn(dA * Iy[8 + k] + dB * Iy[12 + k]);
            }
          } else if (hasB) {
  #pragma unroll
            for (int i = 0; i < 5; i++) {
              const int dA = jv[i] - Iw[8 + i];
              fterms[4 * NB_SIMD + 5 * rowB + i] = __int2float_rn(dA * Ix[8 + i]);
              fterms[FB_WORDS / 2 + 4 * NB_SIMD + 5 * rowB + i] = __int2float_rn(dA * Iy[8 + i]);
            }
          }
          __syncwarp();
          const float acc = run_chain(fterms, lane, 2, NB_SIMD);
          b1f = chain_combine(__shfl_sync(FULL, acc, 0), __shfl_sync(FULL, acc, 1), __shfl_sync(FULL, acc, 2),
                              __shfl_sync(FULL, acc, 3), __shfl_sync(FULL, acc, 4));
          b2f = chain_combine(__shfl_sync(FULL, acc, 5), __shfl_sync(FULL, acc, 6), __shfl_sync(FULL, acc, 7),
                              __shfl_sync(FULL, acc, 8), __shfl_sync(FULL, acc, 9));
        }
      }
      const float b1 = __fmul_rn(b1f, FLT_SCALE);
      const float b2 = __fmul_rn(b2f, FLT_SCALE);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      nx = __fadd_rn(nx, dx);
      ny = __fadd_rn(ny, dy);
      outx = __fadd_rn(nx, half_win);
      outy = __fadd_rn(ny, half_win);
      {
        // OpenCV tests (double)dx*dx + (double)dy*dy <= eps^2.  The float value of that sum is within
        // 3 ulp of it, so the double evaluation is only needed inside a narrow band around eps^2.
        const float s2 = fmaf(dx, dx, dy * dy);
        bool conv;
        if (s2 < eps_lo) conv = true;
        else if (s2 > eps_hi) conv = false;
        else conv = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps_sq;
        if (conv) break;
      }
      // OpenCV: std::abs(delta.x + prevDelta.x) < 0.01 in double; 0.01f is the largest float below 0.01
      if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
        outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f));
        outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
        break;
      }
      pdx = dx;
      pdy = dy;
    }

    // ---- err pass (OpenCV computes it whenever an err array is passed; it can clear status)
    if (st && level == 0) {
      const float fx = __fsub_rn(outx, half_win), fy = __fsub_rn(outy, half_win);
      const int inx = (int)floorf(fx), iny = (int)floorf(fy);
      if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
        st = false;
      } else if (err) {   // callers that do not read err (the reference never does) skip the sum, not the test above
        lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), wt, wb, iw00, iw01, iw10, iw11);
        int bx = inx + PAD_L - tX0, by = iny + PAD_Y - tY0;
        if ((unsigned)bx > 10u || (unsigned)by > 6u) {
          __syncwarp();
          tX0 = (inx + PAD_L - MARGIN) & ~3;
          tY0 = iny + PAD_Y - MARGIN;
          stage_tile(jtile + st_lane, J.img + (size_t)tY0 * pitch + tX0, pitch, lane);
          cp_async_commit();
          cp_async_wait<0>();
          __syncwarp();
          bx = inx + PAD_L - tX0;
          by = MARGIN;
        }
        unsigned se = 0;
        int jv[8];
        const unsigned* jp = jtile + by * TS + (bx >> 2);
        const unsigned shJ = (unsigned)(bx & 3) * 8;
        oct_sample(jp + offA, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) se = __sad(jv[i], Iw[i], se);
        oct_sample(jp + offB, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const bool valid = hasB && (!isq_b || i < 5);
          se += valid ? __sad(jv[i], Iw[8 + i], 0u) : 0u;
        }
        const int tot = __reduce_add_sync(FULL, (int)se);  // <= 441*8160 < 2^24: OpenCV's float sum is exact
        errv = __fdiv_rn((float)(CN * tot), (float)(32 * WIN * CN * WIN));
      }
    }
  }

  if (lane == 0) {
    next_pts[warp] = make_float2(outx, outy);
    status[warp] = st ? 1 : 0;
    if (err) err[warp] = errv;
    if (work) {
      atomicAdd(&work[0], (unsigned long long)n_levels_done);
      atomicAdd(&work[1], (unsigned long long)n_iters_done);
      atomicAdd(&work[2], (unsigned long long)n_slow_a);
      atomicAdd(&work[3], (unsigned long long)n_slow_b);
    }
  }
}

// ------------------------------------------------------------------------------------ 3-channel (BGR) LK
// cv::calcOpticalFlowPyrLK on 3-channel images -- what the reference actually feeds it (imread's default BGR
// output, reference src/keyFrameManagement.cpp:52,64).  OpenCV walks a window row as 63 INTERLEAVED samples
// x = 3*col + ch in steps of 8: the SIMD part is x < 56 (sample x -> float lane x & 3, pairs (x, x + 4) inside a
// step of 8), the scalar chain takes x = 56..62 (column 18 channel 2, columns 19 and 20).  With the planar storage
// of common.cuh that is THREE WARPS per keypoint, one per plane, each running the 1-channel mapping on its plane
// (same octs / tails, same staging, same DP2A samples); x & 3 == (ch - col) & 3, so in an oct (col0 = 0 or 8) the
// sample i of plane ch belongs to chain (ch - i) & 3 -- the per-lane accumulators are the 1-channel ones, only
// their chain index is rotated by the plane.  Of a tail (columns 16..20) the samples with 3*i + ch < 8 are SIMD
// samples (i = 0, 1 and, for ch < 2, i = 2), the rest belong to the scalar chain.  The three warps exchange their
// per-chain integer sums through shared memory (one __syncthreads per exchange, double-buffered) and then all run
// the same float arithmetic on the same numbers, so they take every branch together.  The sequential fall-back
// (a partial sum may reach 2^24) gathers the float terms of all planes in one shared buffer and lanes 0..4 of
// warp 0 add them up in OpenCV's interleaved order.  minEig is normalised by the window AREA (no channel
// factor) and err by area * cn, as in OpenCV.
namespace {
constexpr int C3 = 3;
constexpr int XROW = WIN * C3;              // 63 interleaved samples per window row
constexpr int XSIMD = (XROW / 8) * 8;       // 56
constexpr int PBUF_WORDS = WIN * XROW;      // 1323: one value per window sample, OpenCV's interleaved order
constexpr int XCH_WORDS = 2 * C3 * 32;      // exchange area: [phase][plane][32]
constexpr int PB_REGION = (C3 * PBUF_WORDS > C3 * SCRATCH_WORDS) ? C3 * PBUF_WORDS : C3 * SCRATCH_WORDS;   // 3969: three term
                                            // buffers (one per sum), aliasing the three scratch regions

// sum over the three planes of NV warp-uniform values; pos[k] = slot of value k (a rotation of the chain index)
template <int NV>
__device__ __forceinline__ void plane_combine(int* xch, unsigned& phase, int ch, int lane, int v[NV], const int pos[NV]) {
  int* buf = xch + (phase & 1u) * (C3 * 32);
  phase++;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) buf[ch * 32 + pos[k]] = v[k];
  }
  __syncthreads();
  int mine = 0;
  if (lane < NV) mine = (int)((unsigned)buf[lane] + (unsigned)buf[32 + lane] + (unsigned)buf[64 + lane]);
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = __shfl_sync(FULL, mine, k);
}

// per-lane chain totals in the LOCAL index space q = i & 3 (true chain = (ch - q) & 3): slot A is an oct, slot B an
// oct (b[i] of all 8 samples) or a tail whose first `nsimd` samples are SIMD samples and the others scalar
__device__ __forceinline__ void lane_chains3(const int a[4], const int b[8], bool tail, int nsimd, int c[5]) {
  int sc = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const bool simd_lo = !tail || q < nsimd;
    c[q] = a[q] + (simd_lo ? b[q] : 0) + (tail ? 0 : b[q + 4]);
    sc += (tail && !simd_lo ? b[q] : 0) + (tail ? b[q + 4] : 0);
  }
  c[4] = sc;
}

// One chain of a b sum over the interleaved int32 products P[row][x] of all planes, evaluated by ONE WARP:
// element j = (row r, step t): SIMD chain k < 4 -> P[r][8t+k] + P[r][8t+k+4] (OpenCV's int32 pair), scalar chain ->
// P[r][56+t].  Lane r < 21 walks its row; a warp scan gives the exact prefix in front of every row.  If no partial
// sum and no element reaches 2^24 the float chain is exact and equals the total; otherwise lane 0 adds the float
// terms one by one in OpenCV's order.  (int32 wrap-around cannot hide a violation: elements are < 2^27, so a
// prefix cannot get from (-2^24, 2^24) to an aliasing value without passing through a flagged one.)
__device__ __forceinline__ float chain_b3(const int* P, int k, int lane) {
  int e[7];
  int run = 0;
  const int* row = P + (lane < WIN ? lane : 0) * XROW;
#pragma unroll
  for (int t = 0; t < 7; t++) {
    e[t] = k < 4 ? row[8 * t + k] + row[8 * t + k + 4] : row[XSIMD + t];
    if (lane >= WIN) e[t] = 0;
    run += e[t];
  }
  int incl = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += v;
  }
  int p = incl - run;      // exact prefix in front of this row
  int worst = 0;
#pragma unroll
  for (int t = 0; t < 7; t++) {
    p += e[t];
    worst = max(worst, max(abs(p), abs(e[t])));
  }
  const int total = __shfl_sync(FULL, incl, 31);
  const bool exact = __reduce_max_sync(FULL, min(worst, SAFE_LIMIT)) < SAFE_LIMIT;
  float acc = (float)total;
  if (!exact) {
    acc = 0.f;
    if (lane == 0) {
      for (int r = 0; r < WIN; r++)
        for (int t = 0; t < 7; t++) {
          const int* q = P + r * XROW;
          acc = __fadd_rn(acc, __int2float_rn(k < 4 ? q[8 * t + k] + q[8 * t + k + 4] : q[XSIMD + t]));
        }
    }
    acc = __shfl_sync(FULL, acc, 0);
  }
  return acc;
}
}  // namespace

__global__ void __launch_bounds__(C3 * 32, 5)
lk_kernel_c3(PyrView prev, PyrView next, const float2* __restrict__ prev_pts, int n, float2* __restrict__ next_pts,
             uint8_t* __restrict__ status, float* __restrict__ err, int max_iters, double eps_sq, float eps_lo, float eps_hi,
             float min_eig_thr, unsigned long long* __restrict__ work, const int* __restrict__ n_dev, int skip_mono) {
  if (n_dev) n = min(n, *n_dev);
  if (skip_mono && *prev.mono && *next.mono) return;     // three identical planes in both images: lk_kernel<.., true>
  __shared__ __align__(16) unsigned smem[C3 * TILE_WORDS + PB_REGION + XCH_WORDS];
  const int point = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int ch = threadIdx.x >> 5;            // this warp's plane
  if (point >= n) return;
  unsigned* jtile = smem + ch * TILE_WORDS;
  unsigned* itile = smem + C3 * TILE_WORDS + ch * SCRATCH_WORDS;
  unsigned* dtile = itile + TILE_WORDS;
  unsigned* pbuf = smem + C3 * TILE_WORDS;     // all planes' terms (slow paths; the scratch regions are dead by then)
  int* xch = reinterpret_cast<int*>(smem + C3 * TILE_WORDS + PB_REGION);
  unsigned phase = 0;
  const int st_lane = (lane >> 3) * TS + (lane & 7);
  const float2 pt = prev_pts[point];
  const float half_win = (WIN - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);

  unsigned cfg;
  {
    const int rA = lane >> 1, cA = (lane & 1) * 8;
    const bool q = lane >= 10, hb = lane < 31;
    const int rB = q ? (hb ? lane - 10 : 0) : 16 + (lane >> 1);
    const int cB = q ? 16 : (lane & 1) * 8;
    const unsigned v = (unsigned)(rA * TS + (cA >> 2)) | ((unsigned)(rB * TS + (cB >> 2)) << 8) | ((unsigned)rA << 16) |
                       ((unsigned)rB << 21) | ((unsigned)(cA >> 3) << 26) | ((unsigned)(cB >> 3) << 27) | (q ? 1u << 29 : 0u) |
                       (hb ? 1u << 30 : 0u);
    asm volatile("mov.b32 %0, %1;" : "=r"(cfg) : "r"(v));
  }
  const int nsimd = ch == 2 ? 2 : 3;          // SIMD samples of a tail in this plane (3*i + ch < 8)
  // slot of the local chain q in the exchange rows: the true chain index
  int rot[4];
#pragma unroll
  for (int q = 0; q < 4; q++) rot[q] = (ch - q) & 3;

  float outx = 0.f, outy = 0.f;
  bool st = true;
  float errv = 0.f;
  unsigned int n_levels_done = 0, n_iters_done = 0, n_slow_a = 0, n_slow_b = 0;

  int Iw[16], Ix[16], Iy[16];
  int mpA[4], mB[8];

  const int top = prev.nlevels - 1;
  for (int level = top; level >= 0; level--) {
    const PyrLevelView I = prev.lv[level];
    const PyrLevelView J = next.lv[level];
    const int pitch = I.pitch;
    const uint8_t* Iimg = I.img + (size_t)ch * I.plane;
    const short2* Ider = I.deriv + (size_t)ch * I.plane;
    const uint8_t* Jimg = J.img + (size_t)ch * J.plane;
    const float scale = __int_as_float((127 - level) << 23);   // 2^-level
    float px = __fmul_rn(pt.x, scale), py = __fmul_rn(pt.y, scale);
    float nx, ny;
    if (level == top) {
      nx = px;
      ny = py;
    } else {
      nx = __fmul_rn(outx, 2.f);
      ny = __fmul_rn(outy, 2.f);
    }
    outx = nx;
    outy = ny;

    px = __fsub_rn(px, half_win);
    py = __fsub_rn(py, half_win);
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -WIN || ipx >= I.w || ipy < -WIN || ipy >= I.h) {
      if (level == 0) {
        st = false;
        errv = 0.f;
      }
      continue;
    }
    int iw00, iw01, iw10, iw11;
    unsigned wt, wb;
    lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), wt, wb, iw00, iw01, iw10, iw11);

    nx = __fsub_rn(nx, half_win);
    ny = __fsub_rn(ny, half_win);
    int tX0 = -(1 << 20), tY0 = -(1 << 20);
    __syncthreads();                             // every plane is done with the tiles / term buffer of the previous level
    const int iX = ipx + PAD_L, iY = ipy + PAD_Y;
    {
      stage_tile(itile + st_lane, Iimg + (size_t)iY * pitch + (iX & ~3), pitch, lane);
      {
        const unsigned* dsrc = reinterpret_cast<const unsigned*>(Ider + (size_t)iY * pitch + iX) + (lane >> 3) * pitch + (lane & 7);
        unsigned* ddst = dtile + (lane >> 3) * DS + (lane & 7);
#pragma unroll
        for (int i = 0; i < 6; i++) {
          if (i < 5 || (lane >> 3) < DROWS - 20) {
            cp_async4(ddst, dsrc);
            cp_async4(ddst + 8, dsrc + 8);
            cp_async4(ddst + 16, dsrc + 16);
          }
          dsrc += 4 * pitch;
          ddst += 4 * DS;
        }
      }
      cp_async_commit();
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (!(inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h)) {
        tX0 = (inx + PAD_L - MARGIN) & ~3;
        tY0 = iny + PAD_Y - MARGIN;
        stage_tile(jtile + st_lane, Jimg + (size_t)tY0 * pitch + tX0, pitch, lane);
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
    }

    // ---- window extraction of this plane
    int sums[25];     // A11 | A12 | A22 | C1 | C2, five chains each (local chain order until the exchange)
    {
      const unsigned shI = (unsigned)(iX & 3) * 8;
      oct_sample(itile + offA, shI, wt, wb, Iw);
      oct_sample(itile + offB, shI, wt, wb, Iw + 8);
      oct_deriv(dtile + rowA * DS + colA, iw00, iw01, iw10, iw11, Ix, Iy);
      oct_deriv(dtile + rowB * DS + colB, iw00, iw01, iw10, iw11, Ix + 8, Iy + 8);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const bool valid = hasB && (!isq_b || i < 5);
        if (!valid) { Iw[8 + i] = 0; Ix[8 + i] = 0; Iy[8 + i] = 0; }
      }
      int a11[4] = {0, 0, 0, 0}, a12[4] = {0, 0, 0, 0}, a22[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
      int b11[8], b12[8], b22[8], d1[8], d2[8];
      int mA[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int k = i & 3;
        a11[k] += Ix[i] * Ix[i];
        a12[k] += Ix[i] * Iy[i];
        a22[k] += Iy[i] * Iy[i];
        c1[k] += Iw[i] * Ix[i];
        c2[k] += Iw[i] * Iy[i];
        mA[i] = max(abs(Ix[i]), abs(Iy[i]));
        b11[i] = Ix[8 + i] * Ix[8 + i];
        b12[i] = Ix[8 + i] * Iy[8 + i];
        b22[i] = Iy[8 + i] * Iy[8 + i];
        d1[i] = Iw[8 + i] * Ix[8 + i];
        d2[i] = Iw[8 + i] * Iy[8 + i];
        mB[i] = max(abs(Ix[8 + i]), abs(Iy[8 + i]));
      }
#pragma unroll
      for (int k = 0; k < 4; k++) mpA[k] = max(mA[k], mA[k + 4]);
      lane_chains3(a11, b11, isq_b, nsimd, sums + 0);
      lane_chains3(a12, b12, isq_b, nsimd, sums + 5);
      lane_chains3(a22, b22, isq_b, nsimd, sums + 10);
      lane_chains3(c1, d1, isq_b, nsimd, sums + 15);
      lane_chains3(c2, d2, isq_b, nsimd, sums + 20);
    }
    int pos25[25];
#pragma unroll
    for (int s = 0; s < 5; s++) {
#pragma unroll
      for (int q = 0; q < 4; q++) pos25[s * 5 + q] = s * 5 + rot[q];
      pos25[s * 5 + 4] = s * 5 + 4;
    }
#pragma unroll
    for (int k = 0; k < 25; k++) sums[k] = __reduce_add_sync(FULL, sums[k]);
    // The A11 / A22 chain totals are sums of squares and bound every partial sum of their chain (and of the A12
    // chain: |Ix*Iy| <= max(Ix^2, Iy^2)).  One plane's part (<= 98 * 4080^2 < 2^31) cannot wrap, the sum over the
    // planes could: clamp the parts at 2^24 -- a clamped value only ever feeds the "unsafe" decision.
#pragma unroll
    for (int k = 0; k < 5; k++) {
      sums[k] = min(sums[k], SAFE_LIMIT);
      sums[10 + k] = min(sums[10 + k], SAFE_LIMIT);
    }
    plane_combine<25>(xch, phase, ch, lane, sums, pos25);      // now in TRUE chain order, summed over the planes
    const int* cA11 = sums + 0;
    const int* cA12 = sums + 5;
    const int* cA22 = sums + 10;
    const int* cC1 = sums + 15;
    const int* cC2 = sums + 20;
    bool safeA = true;
#pragma unroll
    for (int k = 0; k < 5; k++) safeA = safeA && cA11[k] < SAFE_LIMIT && cA22[k] < SAFE_LIMIT;
    float A11, A12, A22;
    if (safeA) {
      A11 = chain_combine((float)cA11[0], (float)cA11[1], (float)cA11[2], (float)cA11[3], (float)cA11[4]);
      A12 = chain_combine((float)cA12[0], (float)cA12[1], (float)cA12[2], (float)cA12[3], (float)cA12[4]);
      A22 = chain_combine((float)cA22[0], (float)cA22[1], (float)cA22[2], (float)cA22[3], (float)cA22[4]);
    } else {
      // slow path: the float terms of all planes in OpenCV's interleaved order, one buffer per A sum; warp s adds
      // up sum s (lanes 0..4 = its five chains)
      if (ch == 0) n_slow_a++;
      __syncthreads();      // tiles consumed by every plane
#pragma unroll
      for (int s = 0; s < 3; s++) {
        float* ft = reinterpret_cast<float*>(pbuf) + s * PBUF_WORDS;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int pa = s == 0 ? Ix[i] * Ix[i] : (s == 1 ? Ix[i] * Iy[i] : Iy[i] * Iy[i]);
          ft[rowA * XROW + 3 * (colA + i) + ch] = __int2float_rn(pa);
          const int pb = s == 0 ? Ix[8 + i] * Ix[8 + i] : (s == 1 ? Ix[8 + i] * Iy[8 + i] : Iy[8 + i] * Iy[8 + i]);
          if (hasB && (!isq_b || i < 5)) ft[rowB * XROW + 3 * (colB + i) + ch] = __int2float_rn(pb);
        }
      }
      __syncthreads();
      {
        const float* ft = reinterpret_cast<const float*>(pbuf) + ch * PBUF_WORDS;
        float acc = 0.f;
        if (lane < 4) {
          for (int r = 0; r < WIN; r++)
            for (int x = lane; x < XSIMD; x += 4) acc = __fadd_rn(acc, ft[r * XROW + x]);
        } else if (lane == 4) {
          for (int r = 0; r < WIN; r++)
            for (int x = XSIMD; x < XROW; x++) acc = __fadd_rn(acc, ft[r * XROW + x]);
        }
        if (lane < 5) xch[(phase & 1u) * (C3 * 32) + ch * 32 + lane] = __float_as_int(acc);
      }
      __syncthreads();
      float res[3];
      {
        const int* rb = xch + (phase & 1u) * (C3 * 32);
        phase++;
#pragma unroll
        for (int s = 0; s < 3; s++)
          res[s] = chain_combine(__int_as_float(rb[s * 32 + 0]), __int_as_float(rb[s * 32 + 1]), __int_as_float(rb[s * 32 + 2]),
                                 __int_as_float(rb[s * 32 + 3]), __int_as_float(rb[s * 32 + 4]));
      }
      A11 = res[0];
      A12 = res[1];
      A22 = res[2];
    }
    const int C1tot = (int)((unsigned)cC1[0] + (unsigned)cC1[1] + (unsigned)cC1[2] + (unsigned)cC1[3] + (unsigned)cC1[4]);
    const int C2tot = (int)((unsigned)cC2[0] + (unsigned)cC2[1] + (unsigned)cC2[2] + (unsigned)cC2[3] + (unsigned)cC2[4]);
    int kC1[5], kC2[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { kC1[k] = cC1[k]; kC2[k] = cC2[k]; }
    A11 = __fmul_rn(A11, FLT_SCALE);
    A12 = __fmul_rn(A12, FLT_SCALE);
    A22 = __fmul_rn(A22, FLT_SCALE);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float q = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(q)), (float)(2 * WIN * WIN));
    n_levels_done++;
    cp_async_wait<0>();
    __syncwarp();
    if (min_eig < min_eig_thr || D < 1.1920929e-07f) {
      if (level == 0) st = false;
      continue;
    }
    D = __fdiv_rn(1.f, D);

    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_iters; j++) {
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
        if (level == 0) st = false;
        break;
      }
      lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), wt, wb, iw00, iw01, iw10, iw11);
      int bx = inx + PAD_L - tX0, by = iny + PAD_Y - tY0;
      if ((unsigned)bx > 10u || (unsigned)by > 6u) {   // the window left the staged tile
        __syncwarp();
        tX0 = (inx + PAD_L - MARGIN) & ~3;
        tY0 = iny + PAD_Y - MARGIN;
        stage_tile(jtile + st_lane, Jimg + (size_t)tY0 * pitch + tX0, pitch, lane);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        bx = inx + PAD_L - tX0;
        by = MARGIN;
      }
      const unsigned* jp = jtile + by * TS + (bx >> 2);
      const unsigned shJ = (unsigned)(bx & 3) * 8;
      int a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0}, ua[4];
      int b1[8], b2[8], ub[8];
      int tot[3];
      {
        int jv[8];
        oct_sample(jp + offA, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          a1[i & 3] += jv[i] * Ix[i];
          a2[i & 3] += jv[i] * Iy[i];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) ua[k] = (int)__sad(jv[k + 4], Iw[k + 4], __sad(jv[k], Iw[k], 0u)) * mpA[k];
        oct_sample(jp + offB, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          b1[i] = jv[i] * Ix[8 + i];
          b2[i] = jv[i] * Iy[8 + i];
          ub[i] = (int)__sad(jv[i], Iw[8 + i], 0u) * mB[i];
        }
        const int t1 = ((a1[0] + a1[1]) + (a1[2] + a1[3])) + (((b1[0] + b1[1]) + (b1[2] + b1[3])) + ((b1[4] + b1[5]) + (b1[6] + b1[7])));
        const int t2 = ((a2[0] + a2[1]) + (a2[2] + a2[3])) + (((b2[0] + b2[1]) + (b2[2] + b2[3])) + ((b2[4] + b2[5]) + (b2[6] + b2[7])));
        const int tu = ((ua[0] + ua[1]) + (ua[2] + ua[3])) + (((ub[0] + ub[1]) + (ub[2] + ub[3])) + ((ub[4] + ub[5]) + (ub[6] + ub[7])));
        tot[0] = (int)min(__reduce_add_sync(FULL, (unsigned)min(tu, BOUND_CLAMP)), (unsigned)BOUND_CLAMP);
        tot[1] = __reduce_add_sync(FULL, t1);
        tot[2] = __reduce_add_sync(FULL, t2);
      }
      if (ch == 0) n_iters_done++;
      {
        const int pos3[3] = {0, 1, 2};
        plane_combine<3>(xch, phase, ch, lane, tot, pos3);
      }
      float b1f, b2f;
      const int T1 = (int)((unsigned)tot[1] - (unsigned)C1tot), T2 = (int)((unsigned)tot[2] - (unsigned)C2tot);
      if ((unsigned)tot[0] + (unsigned)max(abs(T1), abs(T2)) < 2u * SAFE_LIMIT) {
        // tier 1: every chain and every step of the final combination is exact (bound U + |T| < 2^25, see lk_kernel)
        b1f = (float)T1;
        b2f = (float)T2;
      } else {
        int cs[15];
        lane_chains3(a1, b1, isq_b, nsimd, cs + 0);
        lane_chains3(a2, b2, isq_b, nsimd, cs + 5);
        lane_chains3(ua, ub, isq_b, nsimd, cs + 10);
        int pos15[15];
#pragma unroll
        for (int s = 0; s < 3; s++) {
#pragma unroll
          for (int q = 0; q < 4; q++) pos15[s * 5 + q] = s * 5 + rot[q];
          pos15[s * 5 + 4] = s * 5 + 4;
        }
#pragma unroll
        for (int k = 0; k < 10; k++) cs[k] = __reduce_add_sync(FULL, cs[k]);
#pragma unroll
        for (int k = 10; k < 15; k++)
          cs[k] = (int)min(__reduce_add_sync(FULL, (unsigned)min(cs[k], BOUND_CLAMP)), (unsigned)BOUND_CLAMP);
        plane_combine<15>(xch, phase, ch, lane, cs, pos15);
        bool safe = true;
        int e1[5], e2[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
          e1[k] = (int)((unsigned)cs[k] - (unsigned)kC1[k]);
          e2[k] = (int)((unsigned)cs[5 + k] - (unsigned)kC2[k]);
          safe = safe && (unsigned)cs[10 + k] + (unsigned)max(abs(e1[k]), abs(e2[k])) < 2u * SAFE_LIMIT;
        }
        if (safe) {
          b1f = chain_combine((float)e1[0], (float)e1[1], (float)e1[2], (float)e1[3], (float)e1[4]);
          b2f = chain_combine((float)e2[0], (float)e2[1], (float)e2[2], (float)e2[3], (float)e2[4]);
        } else {
          // The bound on sum |term| failed: exchange the int32 products of all planes (interleaved order) and
          // evaluate the ten chains exactly, chain c on warp c % 3 (chain_b3: prefix test, sequential fall-back)
          if (ch == 0) n_slow_b++;
          int* p1 = reinterpret_cast<int*>(pbuf);
          int* p2 = p1 + PBUF_WORDS;
          __syncthreads();
          {
            int jv[8];
            oct_sample(jp + offA, shJ, wt, wb, jv);
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int d = jv[i] - Iw[i], o = rowA * XROW + 3 * (colA + i) + ch;
              p1[o] = d * Ix[i];
              p2[o] = d * Iy[i];
            }
            oct_sample(jp + offB, shJ, wt, wb, jv);
#pragma unroll
            for (int i = 0; i < 8; i++)
              if (hasB && (!isq_b || i < 5)) {
                const int d = jv[i] - Iw[8 + i], o = rowB * XROW + 3 * (colB + i) + ch;
                p1[o] = d * Ix[8 + i];
                p2[o] = d * Iy[8 + i];
              }
          }
          __syncthreads();
          int* xb = xch + (phase & 1u) * (C3 * 32);
          for (int c = ch; c < 10; c += C3) {
            const float v = chain_b3(c < 5 ? p1 : p2, c < 5 ? c : c - 5, lane);
            if (lane == 0) xb[c] = __float_as_int(v);
          }
          __syncthreads();
          phase++;
          float resb[2];
#pragma unroll
          for (int s = 0; s < 2; s++)
            resb[s] = chain_combine(__int_as_float(xb[s * 5 + 0]), __int_as_float(xb[s * 5 + 1]), __int_as_float(xb[s * 5 + 2]),
                                    __int_as_float(xb[s * 5 + 3]), __int_as_float(xb[s * 5 + 4]));
          b1f = resb[0];
          b2f = resb[1];
        }
      }
      const float bb1 = __fmul_rn(b1f, FLT_SCALE);
      const float bb2 = __fmul_rn(b2f, FLT_SCALE);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, bb2), __fmul_rn(A22, bb1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, bb1), __fmul_rn(A11, bb2)), D);
      nx = __fadd_rn(nx, dx);
      ny = __fadd_rn(ny, dy);
      outx = __fadd_rn(nx, half_win);
      outy = __fadd_rn(ny, half_win);
      {
        const float s2 = fmaf(dx, dx, dy * dy);
        bool conv;
        if (s2 < eps_lo) conv = true;
        else if (s2 > eps_hi) conv = false;
        else conv = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps_sq;
        if (conv) break;
      }
      if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
        outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f));
        outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
        break;
      }
      pdx = dx;
      pdy = dy;
    }

    // ---- err pass: mean |J - I| / 32 over the window and the channels at the final position
    if (st && level == 0) {
      const float fx = __fsub_rn(outx, half_win), fy = __fsub_rn(outy, half_win);
      const int inx = (int)floorf(fx), iny = (int)floorf(fy);
      if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
        st = false;
      } else if (err) {
        lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), wt, wb, iw00, iw01, iw10, iw11);
        int bx = inx + PAD_L - tX0, by = iny + PAD_Y - tY0;
        if ((unsigned)bx > 10u || (unsigned)by > 6u) {
          __syncwarp();
          tX0 = (inx + PAD_L - MARGIN) & ~3;
          tY0 = iny + PAD_Y - MARGIN;
          stage_tile(jtile + st_lane, Jimg + (size_t)tY0 * pitch + tX0, pitch, lane);
          cp_async_commit();
          cp_async_wait<0>();
          __syncwarp();
          bx = inx + PAD_L - tX0;
          by = MARGIN;
        }
        unsigned se = 0;
        int jv[8];
        const unsigned* jp = jtile + by * TS + (bx >> 2);
        const unsigned shJ = (unsigned)(bx & 3) * 8;
        oct_sample(jp + offA, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) se = __sad(jv[i], Iw[i], se);
        oct_sample(jp + offB, shJ, wt, wb, jv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const bool valid = hasB && (!isq_b || i < 5);
          se += valid ? __sad(jv[i], Iw[8 + i], 0u) : 0u;
        }
        int tot[1] = {__reduce_add_sync(FULL, (int)se)};
        const int pos1[1] = {0};
        plane_combine<1>(xch, phase, ch, lane, tot, pos1);   // <= 1323*8160 < 2^24: OpenCV's float sum is exact
        errv = __fdiv_rn((float)tot[0], (float)(32 * WIN * C3 * WIN));
      }
    }
  }

  if (ch == 0 && lane == 0) {
    next_pts[point] = make_float2(outx, outy);
    status[point] = st ? 1 : 0;
    if (err) err[point] = errv;
    if (work) {
      atomicAdd(&work[0], (unsigned long long)n_levels_done);
      atomicAdd(&work[1], (unsigned long long)n_iters_done);
      atomicAdd(&work[2], (unsigned long long)n_slow_a);
      atomicAdd(&work[3], (unsigned long long)n_slow_b);
    }
  }
}

#undef offA
#undef offB
#undef rowA
#undef rowB
#undef colA
#undef colB
#undef isq_b
#undef hasB
#undef isq
#undef notq

int lk_launch(vo_ctx* c, int slot_prev, int slot_next, const float2* d_prev, int n, float2* d_next, uint8_t* d_status,
              float* d_err) {
  if (n <= 0) return VO_OK;
  VO_TRY(pyr_ensure_deriv(c, slot_prev));
  int max_iters = c->p.lk_max_iters < 0 ? 0 : (c->p.lk_max_iters > 100 ? 100 : c->p.lk_max_iters);
  double eps = c->p.lk_eps < 0 ? 0 : (c->p.lk_eps > 10 ? 10 : c->p.lk_eps);
  eps *= eps;
  const int threads = LK_WARPS * 32;
  const int blocks = div_up(n * 32, threads);
  // OpenCV tests (double)dx*dx + (double)dy*dy <= eps^2; the kernel decides in float outside this band
  const float eps_lo = (float)(eps * (1.0 - 1e-5)), eps_hi = (float)(eps * (1.0 + 1e-5));
  const PyrView pv = pyr_view(c->pyr[slot_prev]), nv = pyr_view(c->pyr[slot_next]);
  const float min_eig = (float)c->p.lk_min_eig;
  {
    LaunchScope ls(c, VO_K_LK);
    if (c->p.channels == 3) {
      // two launches, one of which returns at once: the pair is either "gray read as BGR" (three identical planes in both
      // images, flagged on the device when the frames were split into planes) or genuinely coloured
      lk_kernel<4, true><<<blocks, threads, 0, c->stream>>>(pv, nv, d_prev, n, d_next, d_status, d_err, max_iters, eps, eps_lo,
                                                            eps_hi, min_eig, c->d_lk_work, c->n_dev);
      lk_kernel_c3<<<n, C3 * 32, 0, c->stream>>>(pv, nv, d_prev, n, d_next, d_status, d_err, max_iters, eps, eps_lo, eps_hi,
                                                 min_eig, c->d_lk_work, c->n_dev, 1);
    } else {
      // 128 registers / 16 warps per SM: measured best (168 registers 0.254 ms, 96 registers with spills 0.290 ms)
      lk_kernel<4, false><<<blocks, threads, 0, c->stream>>>(pv, nv, d_prev, n, d_next, d_status, d_err, max_iters, eps, eps_lo,
                                                             eps_hi, min_eig, c->d_lk_work, c->n_dev);
    }
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
