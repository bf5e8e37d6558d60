// lk.cu -- K2: pyramidal Lucas-Kanade tracking, one warp per keypoint, all pyramid levels
// in one launch.  Replaces cv::calcOpticalFlowPyrLK as the reference calls it with all
// defaults (src/tracking.cpp:18 stereo L->R, :52 temporal): 21x21 window, 4 levels,
// 30 iterations / eps 0.01, minEigThreshold 1e-4, 1-channel u8 images.
//
// Arithmetic follows OpenCV's LKTrackerInvoker exactly where it is exact: 14-bit fixed
// point bilinear weights (round-half-even), CV_DESCALE, int16-range patch/derivative
// values, FLT_SCALE = 2^-20, float 2x2 solve with every product and sum rounded
// separately (no FMA contraction), status decided at level 0 only, the oscillation
// half-step rule, the final err pass.  The 441-term window sums are accumulated EXACTLY
// in integers (per-lane int32 partials, warp REDUX on a hi/lo split) and rounded to float
// once; OpenCV accumulates them in float SIMD lanes, which is the only source of
// difference (~1e-5 px typical).  oracle/lk.py is the bit-exact twin.
//
// Mapping: lane l owns window pixels k = l, l+32, ... (<441), i.e. 14 per lane; I-patch,
// Ix, Iy stay in registers across iterations; J is read through the read-only path from
// the padded level (no bounds logic, see common.cuh).
#include "common.cuh"

namespace vo {

constexpr int WIN = LK_WIN;
constexpr int NPIX = WIN * WIN;            // 441
constexpr int PER_LANE = (NPIX + 31) / 32;  // 14
constexpr int W_BITS = 14;

__device__ __forceinline__ long long warp_sum_exact(int v) {
  // exact 64-bit sum of 32 int32 partials (|v| < 2^30) with two REDUX instructions
  const int hi = v >> 12;
  const int lo = v & 4095;
  const int shi = __reduce_add_sync(0xffffffffu, hi);
  const int slo = __reduce_add_sync(0xffffffffu, lo);
  return (long long)shi * 4096 + (long long)slo;
}

__device__ __forceinline__ void lk_weights(float a, float b, int& iw00, int& iw01, int& iw10, int& iw11) {
  const float s = (float)(1 << W_BITS);
  const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
  iw00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), s));
  iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), s));
  iw10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), s));
  iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
}

__global__ void __launch_bounds__(128)
lk_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts, int n, float2* __restrict__ next_pts,
          uint8_t* __restrict__ status, float* __restrict__ err, int max_iters, double eps_sq, float min_eig_thr,
          unsigned long long* __restrict__ work) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float2 pt = prev_pts[warp];
  const float half_win = (WIN - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);

  float outx = 0.f, outy = 0.f;  // nextPts[ptidx] as OpenCV keeps it between levels
  bool st = true;
  float errv = 0.f;
  unsigned int n_levels_done = 0, n_iters_done = 0;

  int Iw[PER_LANE], Ix[PER_LANE], Iy[PER_LANE], off[PER_LANE];

  const int top = prev.nlevels - 1;
  for (int level = top; level >= 0; level--) {
    const PyrLevelView I = prev.lv[level];
    const PyrLevelView J = next.lv[level];
    const int pitch = I.pitch;
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
    lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), iw00, iw01, iw10, iw11);

    // ---- window extraction from the previous image + its Scharr derivative
    int sA11 = 0, sA12 = 0, sA22 = 0;
    {
      const uint8_t* ibase = I.img + (size_t)(ipy + PAD_Y) * pitch + (ipx + PAD_L);
      const short2* dbase = I.deriv + (size_t)(ipy + PAD_Y) * pitch + (ipx + PAD_L);
#pragma unroll
      for (int i = 0; i < PER_LANE; i++) {
        const int k = lane + 32 * i;
        Iw[i] = 0; Ix[i] = 0; Iy[i] = 0; off[i] = 0;
        if (k < NPIX) {
          const int y = k / WIN, x = k - y * WIN;
          const int o = y * pitch + x;
          off[i] = o;
          const uint8_t* s = ibase + o;
          const int ival = ((int)__ldg(s) * iw00 + (int)__ldg(s + 1) * iw01 + (int)__ldg(s + pitch) * iw10 +
                            (int)__ldg(s + pitch + 1) * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
          const short2 d00 = __ldg(dbase + o), d01 = __ldg(dbase + o + 1);
          const short2 d10 = __ldg(dbase + o + pitch), d11 = __ldg(dbase + o + pitch + 1);
          const int ixv = ((int)d00.x * iw00 + (int)d01.x * iw01 + (int)d10.x * iw10 + (int)d11.x * iw11 +
                           (1 << (W_BITS - 1))) >> W_BITS;
          const int iyv = ((int)d00.y * iw00 + (int)d01.y * iw01 + (int)d10.y * iw10 + (int)d11.y * iw11 +
                           (1 << (W_BITS - 1))) >> W_BITS;
          Iw[i] = ival; Ix[i] = ixv; Iy[i] = iyv;
          sA11 += ixv * ixv;
          sA12 += ixv * iyv;
          sA22 += iyv * iyv;
        }
      }
    }
    const float A11 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA11)), FLT_SCALE);
    const float A12 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA12)), FLT_SCALE);
    const float A22 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA22)), FLT_SCALE);
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
      lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), iw00, iw01, iw10, iw11);
      const uint8_t* jbase = J.img + (size_t)(iny + PAD_Y) * pitch + (inx + PAD_L);
      int sb1 = 0, sb2 = 0;
#pragma unroll
      for (int i = 0; i < PER_LANE; i++) {
        const int k = lane + 32 * i;
        if (k < NPIX) {
          const uint8_t* s = jbase + off[i];
          const int jv = ((int)__ldg(s) * iw00 + (int)__ldg(s + 1) * iw01 + (int)__ldg(s + pitch) * iw10 +
                          (int)__ldg(s + pitch + 1) * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
          const int diff = jv - Iw[i];
          sb1 += diff * Ix[i];
          sb2 += diff * Iy[i];
        }
      }
      n_iters_done++;
      const float b1 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb1)), FLT_SCALE);
      const float b2 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb2)), FLT_SCALE);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      nx = __fadd_rn(nx, dx);
      ny = __fadd_rn(ny, dy);
      outx = __fadd_rn(nx, half_win);
      outy = __fadd_rn(ny, half_win);
      if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps_sq) break;
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
      } else {
        lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), iw00, iw01, iw10, iw11);
        const uint8_t* jbase = J.img + (size_t)(iny + PAD_Y) * pitch + (inx + PAD_L);
        int se = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; i++) {
          const int k = lane + 32 * i;
          if (k < NPIX) {
            const uint8_t* s = jbase + off[i];
            const int jv = ((int)__ldg(s) * iw00 + (int)__ldg(s + 1) * iw01 + (int)__ldg(s + pitch) * iw10 +
                            (int)__ldg(s + pitch + 1) * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
            se += abs(jv - Iw[i]);
          }
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

int lk_launch(vo_ctx* c, int slot_prev, int slot_next, const float2* d_prev, int n, float2* d_next, uint8_t* d_status,
              float* d_err) {
  if (n <= 0) return VO_OK;
  VO_TRY(pyr_ensure_deriv(c, slot_prev));
  int max_iters = c->p.lk_max_iters < 0 ? 0 : (c->p.lk_max_iters > 100 ? 100 : c->p.lk_max_iters);
  double eps = c->p.lk_eps < 0 ? 0 : (c->p.lk_eps > 10 ? 10 : c->p.lk_eps);
  eps *= eps;
  const int threads = 128;
  const int blocks = div_up(n * 32, threads);
  {
    LaunchScope ls(c, VO_K_LK);
    lk_kernel<<<blocks, threads, 0, c->stream>>>(pyr_view(c->pyr[slot_prev]), pyr_view(c->pyr[slot_next]), d_prev, n,
                                                 d_next, d_status, d_err, max_iters, eps, (float)c->p.lk_min_eig,
                                                 c->d_lk_work);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
