// lk_v1.cuh -- round-1 LK kernels (exact INTEGER window sums; positions within ~1e-5 px of OpenCV, not
// bit-identical).  Kept only as the A/B baseline of lk.cu (env VO_LK_V1=1) and for the 3-channel path
// until its lane-ordered twin lands.  See lk.cu for the kernel the library runs.
#pragma once
#include "common.cuh"

namespace vo {
namespace v1 {


constexpr int WIN = LK_WIN;
constexpr int SEG = 7;                      // pixels per segment
constexpr int SEGS_PER_ROW = WIN / SEG;     // 3
constexpr int NSEG = WIN * SEGS_PER_ROW;    // 63
constexpr int W_BITS = 14;

__device__ __forceinline__ long long warp_sum_exact(int v) {
  // exact 64-bit sum of 32 int32 partials (|v| < 2^30) with two REDUX instructions
  const int hi = v >> 12;
  const int lo = v & 4095;
  const int shi = __reduce_add_sync(0xffffffffu, hi);
  const int slo = __reduce_add_sync(0xffffffffu, lo);
  return (long long)shi * 4096 + (long long)slo;
}

__device__ __forceinline__ void lk_weights(float a, float b, unsigned& wt, unsigned& wb, int& iw00, int& iw01,
                                           int& iw10, int& iw11) {
  const float s = (float)(1 << W_BITS);
  const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
  iw00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), s));
  iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), s));
  iw10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), s));
  iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
  wt = ((unsigned)iw00 & 0xffffu) | ((unsigned)iw01 << 16);   // weights are in [-1, 2^14]: two s16 per register
  wb = ((unsigned)iw10 & 0xffffu) | ((unsigned)iw11 << 16);
}

// d = c + a.s16[0]*b.u8[0] + a.s16[1]*b.u8[1]  (signed 16-bit weights: iw11 = 2^14 - the other
// three can be -1; unsigned 8-bit pixels).  Plain (non-volatile) asm so that ptxas may schedule it.
__device__ __forceinline__ int dp2a_w(unsigned w, unsigned px, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c));
  return d;
}
// same with bytes 2,3 of px
__device__ __forceinline__ int dp2a_w_hi(unsigned w, unsigned px, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c));
  return d;
}

constexpr int PROWS = WIN + 1;   // 22 source rows
#ifndef LK_OPT_HI
#define LK_OPT_HI 1
#endif
#ifndef LK_OPT_DENSE
#define LK_OPT_DENSE 1
#endif
#ifndef LK_OPT_HOIST
#define LK_OPT_HOIST 1
#endif
constexpr int PS = LK_OPT_DENSE ? 7 : 9;            // tile row stride in 32-bit words (dense: conflict-free stores, 2-way loads)
constexpr int DS = LK_OPT_DENSE ? 22 : 23;           // derivative tile row stride in short2 (dense, same reasoning)
constexpr int TILE_WORDS = PROWS * PS;          // 198
constexpr int DTILE_WORDS = PROWS * DS;         // 506
constexpr int WARP_SMEM_WORDS = TILE_WORDS + DTILE_WORDS;
constexpr int LK_WARPS = 4;

// Stage the 22 x 28-byte patch whose (unaligned) origin is `a0` into `tile`; returns the byte
// offset (0..3) of the origin inside the first staged word.  lane_off = (lane/8)*(pitch/4) +
// lane%8 and step = 4*(pitch/4) are per-level constants, so each load is one 64-bit pointer bump.
// `staged` (optional, warp-uniform) remembers the word-aligned origin that is in the tile: an LK iteration
// usually moves the window by a fraction of a pixel, so the next iteration's patch is the one already staged
// (same rows, same first word, only the byte offset and the bilinear weights change) and the six loads, six
// stores and the exposed L2 latency of re-staging it are skipped.
__device__ __forceinline__ unsigned stage_patch(const uint8_t* a0, int lane_off, int step, unsigned* tile_lane,
                                                bool col_ok, bool last_ok, uintptr_t* staged = nullptr) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(a0);
  if (staged) {
    if (*staged == (a & ~uintptr_t(3))) return (unsigned)(a & 3);
    *staged = a & ~uintptr_t(3);
  }
  const unsigned* w = reinterpret_cast<const unsigned*>(a & ~uintptr_t(3)) + lane_off;
  __syncwarp();   // everyone is done reading the previous tile
  unsigned v[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    v[i] = 0;
    if (col_ok && (i < 5 || last_ok)) v[i] = __ldg(w);
    w += step;
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
    if (col_ok && (i < 5 || last_ok)) tile_lane[i * 4 * PS] = v[i];
  __syncwarp();
  return (unsigned)(a & 3);
}

// bilinear samples (x = 0..6) of the segment at (row, byte offset bo) of a staged tile
__device__ __forceinline__ void seg_bilinear(const unsigned* tile, int row, unsigned bo, unsigned wt, unsigned wb,
                                             int out[SEG]) {
  const unsigned* p = tile + row * PS + (bo >> 2);
  const unsigned sh = (bo & 3) * 8;
  const unsigned t0 = __funnelshift_r(p[0], p[1], sh), t1 = __funnelshift_r(p[1], p[2], sh);
  const unsigned b0 = __funnelshift_r(p[PS], p[PS + 1], sh), b1 = __funnelshift_r(p[PS + 1], p[PS + 2], sh);
#if !LK_OPT_HI
  const unsigned tp[SEG] = {t0, t0 >> 8, t0 >> 16, __funnelshift_r(t0, t1, 24), t1, t1 >> 8, t1 >> 16};
  const unsigned bp[SEG] = {b0, b0 >> 8, b0 >> 16, __funnelshift_r(b0, b1, 24), b1, b1 >> 8, b1 >> 16};
#pragma unroll
  for (int x = 0; x < SEG; x++)
    out[x] = dp2a_w(wb, bp[x], dp2a_w(wt, tp[x], 1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
#else
  // pixel pairs (x, x+1): bytes (0,1),(2,3) of t0 / t1 via dp2a.lo/.hi; the odd ones from the
  // registers shifted by one byte -- 2 extra shifts per row instead of 5
  const unsigned tu = __funnelshift_r(t0, t1, 8), tv = t1 >> 8;
  const unsigned bu = __funnelshift_r(b0, b1, 8), bv = b1 >> 8;
  const int rc = 1 << (W_BITS - 5 - 1);
  out[0] = dp2a_w(wb, b0, dp2a_w(wt, t0, rc)) >> (W_BITS - 5);
  out[1] = dp2a_w(wb, bu, dp2a_w(wt, tu, rc)) >> (W_BITS - 5);
  out[2] = dp2a_w_hi(wb, b0, dp2a_w_hi(wt, t0, rc)) >> (W_BITS - 5);
  out[3] = dp2a_w_hi(wb, bu, dp2a_w_hi(wt, tu, rc)) >> (W_BITS - 5);
  out[4] = dp2a_w(wb, b1, dp2a_w(wt, t1, rc)) >> (W_BITS - 5);
  out[5] = dp2a_w(wb, bv, dp2a_w(wt, tv, rc)) >> (W_BITS - 5);
  out[6] = dp2a_w_hi(wb, b1, dp2a_w_hi(wt, t1, rc)) >> (W_BITS - 5);
#endif
}

#ifndef LK_MINBLOCKS
#define LK_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(128)
lk_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts, int n, float2* __restrict__ next_pts,
          uint8_t* __restrict__ status, float* __restrict__ err, int max_iters, double eps_sq, float min_eig_thr,
          unsigned long long* __restrict__ work, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  __shared__ unsigned smem[LK_WARPS * WARP_SMEM_WORDS];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  unsigned* tile = smem + (threadIdx.x >> 5) * WARP_SMEM_WORDS;
  unsigned* dtile = tile + TILE_WORDS;
  // staging role of this lane: row (lane/8) + 4i, word lane%8 (< 7) of the 22 x 7-word tile
  unsigned* tile_lane = tile + (lane >> 3) * PS + (lane & 7);
  const bool st_col_ok = (lane & 7) < 7, st_last_ok = (lane >> 3) < PROWS - 20;
  const float2 pt = prev_pts[warp];
  const float half_win = (WIN - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);
  const float eps_lo = (float)(eps_sq * (1.0 - 1e-5)), eps_hi = (float)(eps_sq * (1.0 + 1e-5));

  // this lane's two segments (fixed for the whole kernel)
  const int rowA = lane / SEGS_PER_ROW, colA = (lane - rowA * SEGS_PER_ROW) * SEG;
  const int sB = lane + 32;
  const bool hasB = sB < NSEG;
  const int rowB = hasB ? sB / SEGS_PER_ROW : 0, colB = hasB ? (sB - rowB * SEGS_PER_ROW) * SEG : 0;

  float outx = 0.f, outy = 0.f;  // nextPts[ptidx] as OpenCV keeps it between levels
  bool st = true;
  float errv = 0.f;
  unsigned int n_levels_done = 0, n_iters_done = 0;

  int Iw[2 * SEG], Ix[2 * SEG], Iy[2 * SEG];

  const int top = prev.nlevels - 1;
  for (int level = top; level >= 0; level--) {
    const PyrLevelView I = prev.lv[level];
    const PyrLevelView J = next.lv[level];
    const int pitch = I.pitch;
    const int st_step = pitch;                              // 4 rows, in words: 4 * (pitch / 4)
    const int st_off = (lane >> 3) * (pitch >> 2) + (lane & 7);
    const float scale = 1.f / (float)(1 << level);
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
    int ipx = (int)floorf(px), ipy = (int)floorf(py);
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

    // ---- window extraction from the previous image + its Scharr derivative
    int sA11 = 0, sA12 = 0, sA22 = 0, sC1 = 0, sC2 = 0;
    uintptr_t staged = 0;     // the I patch below goes through the tile unconditionally; J patches are cached
    {
      const size_t o0 = (size_t)(ipy + PAD_Y) * pitch + (ipx + PAD_L);
      const unsigned sh = stage_patch(I.img + o0, st_off, st_step, tile_lane, st_col_ok, st_last_ok);
      // derivative patch: 484 short2, row-coalesced (lane -> consecutive elements)
      {
        const unsigned* dsrc = reinterpret_cast<const unsigned*>(I.deriv + o0);
        int r = 0, x = lane;
        if (x >= PROWS) { x -= PROWS; r = 1; }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (r < PROWS) dtile[r * DS + x] = __ldg(dsrc + (size_t)r * pitch + x);
          x += 32 - PROWS; r += 1;                       // advance by 32 elements of 22-wide rows
          if (x >= PROWS) { x -= PROWS; r += 1; }
        }
        __syncwarp();
      }
      seg_bilinear(tile, rowA, colA + sh, wt, wb, Iw);
      if (hasB) seg_bilinear(tile, rowB, colB + sh, wt, wb, Iw + SEG);
#pragma unroll
      for (int sgi = 0; sgi < 2; sgi++) {
        if (sgi == 1 && !hasB) {
#pragma unroll
          for (int x = 0; x < SEG; x++) { Iw[SEG + x] = 0; Ix[SEG + x] = 0; Iy[SEG + x] = 0; }
          break;
        }
        const unsigned* d = dtile + (sgi ? rowB : rowA) * DS + (sgi ? colB : colA);
        unsigned top_[SEG + 1], bot_[SEG + 1];
#pragma unroll
        for (int x = 0; x <= SEG; x++) {
          top_[x] = d[x];
          bot_[x] = d[DS + x];
        }
#pragma unroll
        for (int x = 0; x < SEG; x++) {
          // short2 packed in a word: .x = low half (dx), .y = high half (dy)
          const int ixv = ((int)(short)(top_[x] & 0xffff) * iw00 + (int)(short)(top_[x + 1] & 0xffff) * iw01 +
                           (int)(short)(bot_[x] & 0xffff) * iw10 + (int)(short)(bot_[x + 1] & 0xffff) * iw11 +
                           (1 << (W_BITS - 1))) >> W_BITS;
          const int iyv = (((int)top_[x] >> 16) * iw00 + ((int)top_[x + 1] >> 16) * iw01 + ((int)bot_[x] >> 16) * iw10 +
                           ((int)bot_[x + 1] >> 16) * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
          Ix[sgi * SEG + x] = ixv;
          Iy[sgi * SEG + x] = iyv;
          sA11 += ixv * ixv;
          sA12 += ixv * iyv;
          sA22 += iyv * iyv;
          sC1 += Iw[sgi * SEG + x] * ixv;
          sC2 += Iw[sgi * SEG + x] * iyv;
        }
      }
    }
    const float A11 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA11)), FLT_SCALE);
    const float A12 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA12)), FLT_SCALE);
    const float A22 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA22)), FLT_SCALE);
    // sum(diff*Ix) = sum(J*Ix) - sum(I*Ix): the second term is constant over the iterations
    const long long C1 = LK_OPT_HOIST ? warp_sum_exact(sC1) : 0, C2 = LK_OPT_HOIST ? warp_sum_exact(sC2) : 0;
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float q = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(q)), (float)(2 * WIN * WIN));
    n_levels_done++;
    if (min_eig < min_eig_thr || D < 1.1920929e-07f) {
      if (level == 0) st = false;
      continue;
    }
    D = __fdiv_rn(1.f, D);

    nx = __fsub_rn(nx, half_win);
    ny = __fsub_rn(ny, half_win);
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_iters; j++) {
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
        if (level == 0) st = false;
        break;
      }
      lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), wt, wb, iw00, iw01, iw10, iw11);
      const unsigned sh = stage_patch(J.img + ((iny + PAD_Y) * pitch + (inx + PAD_L)), st_off, st_step, tile_lane,
                                      st_col_ok, st_last_ok, &staged);
      int sb1 = 0, sb2 = 0;
      {
        // four independent accumulators per sum (ncu: 29 % of the stalls were `wait`, i.e. the 14-deep
        // dependent IMAD chains of a single accumulator with only 3 warps per scheduler to hide them)
        int p1[4] = {0, 0, 0, 0}, p2[4] = {0, 0, 0, 0};
        int jv[SEG];
        seg_bilinear(tile, rowA, colA + sh, wt, wb, jv);
#pragma unroll
        for (int x = 0; x < SEG; x++) {
          const int dj = LK_OPT_HOIST ? jv[x] : jv[x] - Iw[x];
          p1[x & 1] += dj * Ix[x];
          p2[x & 1] += dj * Iy[x];
        }
        if (hasB) {
          seg_bilinear(tile, rowB, colB + sh, wt, wb, jv);
#pragma unroll
          for (int x = 0; x < SEG; x++) {
            const int dj = LK_OPT_HOIST ? jv[x] : jv[x] - Iw[SEG + x];
            p1[2 + (x & 1)] += dj * Ix[SEG + x];
            p2[2 + (x & 1)] += dj * Iy[SEG + x];
          }
        }
        sb1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
        sb2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
      }
      n_iters_done++;
      const float b1 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb1) - C1), FLT_SCALE);
      const float b2 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb2) - C2), FLT_SCALE);
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
      if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
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
        const unsigned sh = stage_patch(J.img + ((iny + PAD_Y) * pitch + (inx + PAD_L)), st_off, st_step, tile_lane,
                                      st_col_ok, st_last_ok, &staged);
        int se = 0;
        int jv[SEG];
        seg_bilinear(tile, rowA, colA + sh, wt, wb, jv);
#pragma unroll
        for (int x = 0; x < SEG; x++) se += abs(jv[x] - Iw[x]);
        if (hasB) {
          seg_bilinear(tile, rowB, colB + sh, wt, wb, jv);
#pragma unroll
          for (int x = 0; x < SEG; x++) se += abs(jv[x] - Iw[SEG + x]);
        }
        const int tot = __reduce_add_sync(0xffffffffu, se);  // <= 441*8160 fits int32
        errv = __fmul_rn((float)tot, 1.f / (32 * WIN * WIN));
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
    }
  }
}

// ------------------------------------------------------------------------------------ 3-channel (BGR) LK
// cv::calcOpticalFlowPyrLK on 3-channel images -- what the reference actually feeds it (imread's default
// BGR output, reference src/keyFrameManagement.cpp:52,64).  OpenCV walks the window as 21 rows x 63
// interleaved samples (neighbour = +cn): every sum runs over the three channels of the window, the
// bilinear samples and derivatives are per channel.  With planar storage (common.cuh) that is the
// 1-channel computation repeated per plane with ONE set of sums: same segment mapping, same staging,
// same DP2A samples.  Ix/Iy of the 3 x 14 samples a lane owns are kept packed (s16 | s16 << 16) in 42
// registers; I is not kept (the sum of I*Ix is hoisted) and is re-sampled once for the err pass.
// minEig is normalised by the window AREA (no channel factor) and err by area * cn, as in OpenCV.
__device__ __forceinline__ void lk_deriv_seg(const unsigned* dtile, int row, int col, int iw00, int iw01, int iw10, int iw11,
                                             const int* Iw, unsigned* ixy, int& sA11, int& sA12, int& sA22, int& sC1,
                                             int& sC2) {
  const unsigned* d = dtile + row * DS + col;
  unsigned top_[SEG + 1], bot_[SEG + 1];
#pragma unroll
  for (int x = 0; x <= SEG; x++) {
    top_[x] = d[x];
    bot_[x] = d[DS + x];
  }
#pragma unroll
  for (int x = 0; x < SEG; x++) {
    const int ixv = ((int)(short)(top_[x] & 0xffff) * iw00 + (int)(short)(top_[x + 1] & 0xffff) * iw01 +
                     (int)(short)(bot_[x] & 0xffff) * iw10 + (int)(short)(bot_[x + 1] & 0xffff) * iw11 + (1 << (W_BITS - 1))) >>
                    W_BITS;
    const int iyv = (((int)top_[x] >> 16) * iw00 + ((int)top_[x + 1] >> 16) * iw01 + ((int)bot_[x] >> 16) * iw10 +
                     ((int)bot_[x + 1] >> 16) * iw11 + (1 << (W_BITS - 1))) >>
                    W_BITS;
    ixy[x] = ((unsigned)ixv & 0xffffu) | ((unsigned)iyv << 16);
    sA11 += ixv * ixv;
    sA12 += ixv * iyv;
    sA22 += iyv * iyv;
    sC1 += Iw[x] * ixv;
    sC2 += Iw[x] * iyv;
  }
}

__global__ void __launch_bounds__(128)
lk_kernel_c3(PyrView prev, PyrView next, const float2* __restrict__ prev_pts, int n, float2* __restrict__ next_pts,
             uint8_t* __restrict__ status, float* __restrict__ err, int max_iters, double eps_sq, float min_eig_thr,
             unsigned long long* __restrict__ work, const int* __restrict__ n_dev) {
  constexpr int CN = 3;
  if (n_dev) n = min(n, *n_dev);
  // one image tile per channel (so that the J patches of an iteration can stay staged, see stage_patch)
  // + one derivative tile
  constexpr int WARP_WORDS_C3 = CN * TILE_WORDS + DTILE_WORDS;
  __shared__ unsigned smem[LK_WARPS * WARP_WORDS_C3];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  unsigned* tile0 = smem + (threadIdx.x >> 5) * WARP_WORDS_C3;
  unsigned* dtile = tile0 + CN * TILE_WORDS;
  unsigned* tile_lane0 = tile0 + (lane >> 3) * PS + (lane & 7);
  const bool st_col_ok = (lane & 7) < 7, st_last_ok = (lane >> 3) < PROWS - 20;
  const float2 pt = prev_pts[warp];
  const float half_win = (WIN - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);
  const float eps_lo = (float)(eps_sq * (1.0 - 1e-5)), eps_hi = (float)(eps_sq * (1.0 + 1e-5));

  const int rowA = lane / SEGS_PER_ROW, colA = (lane - rowA * SEGS_PER_ROW) * SEG;
  const int sB = lane + 32;
  const bool hasB = sB < NSEG;
  const int rowB = hasB ? sB / SEGS_PER_ROW : 0, colB = hasB ? (sB - rowB * SEGS_PER_ROW) * SEG : 0;

  float outx = 0.f, outy = 0.f;
  bool st = true;
  float errv = 0.f;
  unsigned int n_levels_done = 0, n_iters_done = 0;

  unsigned ixy[CN][2 * SEG];   // packed (Ix, Iy) of this lane's samples, per plane

  const int top = prev.nlevels - 1;
  for (int level = top; level >= 0; level--) {
    const PyrLevelView I = prev.lv[level];
    const PyrLevelView J = next.lv[level];
    const int pitch = I.pitch;
    const int st_step = pitch;
    const int st_off = (lane >> 3) * (pitch >> 2) + (lane & 7);
    const float scale = 1.f / (float)(1 << level);
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
    const unsigned wtI = wt, wbI = wb;
    const size_t o0 = (size_t)(ipy + PAD_Y) * pitch + (ipx + PAD_L);

    // ---- window extraction from the previous image + its Scharr derivative, all planes
    int sA11 = 0, sA12 = 0, sA22 = 0, sC1 = 0, sC2 = 0;
    uintptr_t staged[CN] = {0, 0, 0};
#pragma unroll
    for (int ch = 0; ch < CN; ch++) {
      unsigned* tile = tile0 + ch * TILE_WORDS;
      unsigned* tile_lane = tile_lane0 + ch * TILE_WORDS;
      const unsigned sh = stage_patch(I.img + (size_t)ch * I.plane + o0, st_off, st_step, tile_lane, st_col_ok, st_last_ok);
      {
        const unsigned* dsrc = reinterpret_cast<const unsigned*>(I.deriv + (size_t)ch * I.plane + o0);
        int r = 0, x = lane;
        if (x >= PROWS) { x -= PROWS; r = 1; }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (r < PROWS) dtile[r * DS + x] = __ldg(dsrc + (size_t)r * pitch + x);
          x += 32 - PROWS; r += 1;
          if (x >= PROWS) { x -= PROWS; r += 1; }
        }
        __syncwarp();
      }
      int Iw[SEG];
      seg_bilinear(tile, rowA, colA + sh, wt, wb, Iw);
      lk_deriv_seg(dtile, rowA, colA, iw00, iw01, iw10, iw11, Iw, ixy[ch], sA11, sA12, sA22, sC1, sC2);
      if (hasB) {
        seg_bilinear(tile, rowB, colB + sh, wt, wb, Iw);
        lk_deriv_seg(dtile, rowB, colB, iw00, iw01, iw10, iw11, Iw, ixy[ch] + SEG, sA11, sA12, sA22, sC1, sC2);
      } else {
#pragma unroll
        for (int x = 0; x < SEG; x++) ixy[ch][SEG + x] = 0;
      }
    }
    const float A11 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA11)), FLT_SCALE);
    const float A12 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA12)), FLT_SCALE);
    const float A22 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA22)), FLT_SCALE);
    const long long C1 = warp_sum_exact(sC1), C2 = warp_sum_exact(sC2);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float q = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(q)), (float)(2 * WIN * WIN));
    n_levels_done++;
    if (min_eig < min_eig_thr || D < 1.1920929e-07f) {
      if (level == 0) st = false;
      continue;
    }
    D = __fdiv_rn(1.f, D);

    nx = __fsub_rn(nx, half_win);
    ny = __fsub_rn(ny, half_win);
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_iters; j++) {
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
        if (level == 0) st = false;
        break;
      }
      lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), wt, wb, iw00, iw01, iw10, iw11);
      const size_t oj = (size_t)(iny + PAD_Y) * pitch + (inx + PAD_L);
      int sb1 = 0, sb2 = 0;
#pragma unroll
      for (int ch = 0; ch < CN; ch++) {
        unsigned* tile = tile0 + ch * TILE_WORDS;
        const unsigned sh = stage_patch(J.img + (size_t)ch * J.plane + oj, st_off, st_step, tile_lane0 + ch * TILE_WORDS,
                                        st_col_ok, st_last_ok, &staged[ch]);
        int jv[SEG];
        seg_bilinear(tile, rowA, colA + sh, wt, wb, jv);
#pragma unroll
        for (int x = 0; x < SEG; x++) {
          sb1 += jv[x] * (int)(short)(ixy[ch][x] & 0xffff);
          sb2 += jv[x] * ((int)ixy[ch][x] >> 16);
        }
        if (hasB) {
          seg_bilinear(tile, rowB, colB + sh, wt, wb, jv);
#pragma unroll
          for (int x = 0; x < SEG; x++) {
            sb1 += jv[x] * (int)(short)(ixy[ch][SEG + x] & 0xffff);
            sb2 += jv[x] * ((int)ixy[ch][SEG + x] >> 16);
          }
        }
      }
      n_iters_done++;
      const float b1 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb1) - C1), FLT_SCALE);
      const float b2 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb2) - C2), FLT_SCALE);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
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
      if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
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
        const size_t oj = (size_t)(iny + PAD_Y) * pitch + (inx + PAD_L);
        int se = 0;
#pragma unroll
        for (int ch = 0; ch < CN; ch++) {
          int iv[2 * SEG], jv[SEG];
          unsigned* tile = tile0 + ch * TILE_WORDS;
          unsigned* tile_lane = tile_lane0 + ch * TILE_WORDS;
          const unsigned shI = stage_patch(I.img + (size_t)ch * I.plane + o0, st_off, st_step, tile_lane, st_col_ok, st_last_ok);
          seg_bilinear(tile, rowA, colA + shI, wtI, wbI, iv);
          if (hasB) seg_bilinear(tile, rowB, colB + shI, wtI, wbI, iv + SEG);
          const unsigned sh = stage_patch(J.img + (size_t)ch * J.plane + oj, st_off, st_step, tile_lane, st_col_ok, st_last_ok);
          seg_bilinear(tile, rowA, colA + sh, wt, wb, jv);
#pragma unroll
          for (int x = 0; x < SEG; x++) se += abs(jv[x] - iv[x]);
          if (hasB) {
            seg_bilinear(tile, rowB, colB + sh, wt, wb, jv);
#pragma unroll
            for (int x = 0; x < SEG; x++) se += abs(jv[x] - iv[SEG + x]);
          }
        }
        const int tot = __reduce_add_sync(0xffffffffu, se);  // <= 3*441*8160 fits int32
        errv = __fmul_rn((float)tot, 1.f / (32 * WIN * CN * WIN));
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
    }
  }
}


}  // namespace v1
}  // namespace vo
