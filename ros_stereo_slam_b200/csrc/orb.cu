// orb.cu -- ORB for the loop detector (reference src/optimizationStuff.cpp:49-56: ORB::create()->detectAndCompute feeds
// DBoW2).  SURVEY.md section 8(f)-2.  Every stage is bit-identical to cv2 4.13.0 (oracle/orb.py is the stage-by-stage
// restatement): the INTER_LINEAR_EXACT pyramid, FAST-9/16 with suppression, the Harris ranking response, IC_Angle,
// the smoothing ORB really applies, the rBRIEF descriptors (test-pair table recovered from cv2 itself, orb_pattern.h).
// vo_orb_detect_and_compute runs them for all levels at once with OpenCV's selection rules (quota, border filter,
// retainBest with ties) on the device: one host synchronisation per frame.
//
//   K_rows   the smoothing ORB applies before sampling is NOT OpenCV's fixed-point Gaussian: the pyramid level is a
//            sub-matrix, for which GaussianBlur falls back to the generic float separable filter.  Row pass:
//            s = k0 * p(x-3), then s = fma(k_j, p(x-3+j), s), j = 1..6, BORDER_REFLECT_101 -> float plane
//   K_cols   t = k3 * s(y), then t = fma(k_{3+j}, s(y+j) + s(y-j), t), j = 1..3; rint, saturate -> u8
//            (this evaluation order is the one cv2's AVX2/FMA build uses: zero differing pixels on whole frames)
//   K_fast   cv::FAST (TYPE_9_16), the detector ORB runs on every level: a thread per pixel gathers the 16 circle
//            differences, the corner score (cornerScore<16>) is the best over the 16 arcs of 9 of the smallest margin,
//            minus 1; non-maximum suppression (strictly greater than the 8 neighbours) and a raster-order compaction
//            (CUB DeviceSelect) give cv::FAST's keypoint list, order included
//   K_angle  IC_Angle on the UNSMOOTHED level: a warp per keypoint, lane = column u of the circular patch (radius 15,
//            row half-widths from OpenCV's u_max table), integer moments m_10 / m_01 reduced by shuffles, then
//            cv::fastAtan2 (7th-order polynomial, float, no contraction) -> degrees
//   K_desc   a thread per (keypoint, descriptor byte): angle in degrees -> (float)cos/sin of the double angle, the 16
//            test points of the byte rotated in float without contraction, cvRound, 8 comparisons
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <functional>
#include <numeric>

#include "common.cuh"
#include "orb_pattern.h"

namespace vo {

struct Orb {
  int w = 0, h = 0, cap = 0;
  uint8_t* img = nullptr;
  float* rowf = nullptr;
  uint8_t* sm = nullptr;
  float* xy = nullptr;
  float* ang = nullptr;
  uint8_t* desc = nullptr;
  int* score = nullptr;            // FAST corner scores [h][w]
  uint8_t* flag = nullptr;
  int* sel = nullptr;              // selected flat pixel indices
  int* d_n = nullptr;
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  size_t px_cap = 0;               // pixels the FAST buffers are sized for
  uint8_t* pyr = nullptr;          // pyramid levels 1.. of vo_orb_detect_and_compute, packed
  size_t pyr_bytes = 0;
  int* coef = nullptr;             // resize offsets / weights: ox, cx, oy, cy
  int coef_cap = 0;
  uint8_t* h_pin = nullptr;        // pinned host scratch of vo_orb_detect_and_compute: image | header | packed results
  size_t h_pin_bytes = 0;
  // vo_orb_detect_and_compute, all levels at once: per-pixel buffers over the concatenated levels, packed results
  size_t all_px = 0;
  int* a_score = nullptr;
  uint8_t* a_flag = nullptr;
  int* a_sel = nullptr;
  float* a_rowf = nullptr;
  uint8_t* a_sm = nullptr;
  void* a_cub = nullptr;
  size_t a_cub_bytes = 0;
  int n_cand = 0;
  float2* c_xy = nullptr;          // per-level candidate slices: positions ...
  float* c_resp = nullptr;         // ... and Harris responses
  int out_cap = 0;
  float2* o_xy = nullptr;          // packed over the levels: position * level scale, response, angle, descriptor
  float* o_resp = nullptr;
  float* o_ang = nullptr;
  uint8_t* o_desc = nullptr;
  int* d_hdr = nullptr;            // [0..7] keypoints per level, [8..15] after the first selection, [16] FAST corners
};

__constant__ signed char c_orb_pattern[256][4];
__constant__ float c_orb_gauss[4];

void orb_free(vo_ctx* c) {
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  if (!o) return;
  void* dev[] = {o->img, o->rowf, o->sm, o->xy, o->ang, o->desc, o->score, o->flag, o->sel, o->d_n, o->cub_tmp, o->pyr, o->coef,
                 o->a_score, o->a_flag, o->a_sel, o->a_rowf, o->a_sm, o->a_cub, o->c_xy, o->c_resp, o->o_xy, o->o_resp, o->o_ang,
                 o->o_desc, o->d_hdr};
  for (void* p : dev) cudaFree(p);
  cudaFreeHost(o->h_pin);
  delete o;
  c->orb = nullptr;
}

__device__ __forceinline__ int orb_reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void orb_smooth_rows_kernel(const uint8_t* __restrict__ img, int w, int h, float* __restrict__ rowf) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* r = img + (size_t)y * w;
  float s = __fmul_rn(c_orb_gauss[0], (float)r[orb_reflect101(x - 3, w)]);
#pragma unroll
  for (int j = 1; j < 7; j++) s = __fmaf_rn(c_orb_gauss[j < 4 ? j : 6 - j], (float)r[orb_reflect101(x - 3 + j, w)], s);
  rowf[(size_t)y * w + x] = s;
}

__global__ void orb_smooth_cols_kernel(const float* __restrict__ rowf, int w, int h, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  float t = __fmul_rn(c_orb_gauss[3], rowf[(size_t)y * w + x]);
#pragma unroll
  for (int j = 1; j < 4; j++) {
    const float a = rowf[(size_t)orb_reflect101(y + j, h) * w + x], b = rowf[(size_t)orb_reflect101(y - j, h) * w + x];
    t = __fmaf_rn(c_orb_gauss[3 - j], __fadd_rn(a, b), t);
  }
  const int v = __float2int_rn(t);
  out[(size_t)y * w + x] = (uint8_t)min(max(v, 0), 255);
}

// cv::FAST, patternSize 16: circle offsets in OpenCV's order
__constant__ int c_fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// 9 contiguous set bits in a 16-bit circular mask
__device__ __forceinline__ bool fast_has9(unsigned m) {
  m |= m << 16;
  unsigned r = m & (m >> 1);      // runs of 2
  r &= r >> 2;                    // runs of 4
  r &= r >> 4;                    // runs of 8
  r &= m >> 8;                    // runs of 9
  return (r & 0xFFFFu) != 0u;
}

// score map: cornerScore<16> for corners (>= threshold >= 1), 0 otherwise
__global__ void fast_score_kernel(const uint8_t* __restrict__ img, int w, int h, int threshold, int* __restrict__ score) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  int sc = 0;
  if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
    const uint8_t* p = img + (size_t)y * w + x;
    const int v = p[0];
    int pv[16];
    unsigned dark = 0u, bright = 0u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      pv[k] = p[c_fast_dy[k] * w + c_fast_dx[k]];
      dark |= (unsigned)(pv[k] < v - threshold) << k;
      bright |= (unsigned)(pv[k] > v + threshold) << k;
    }
    if (fast_has9(dark) || fast_has9(bright)) {
      // the largest margin m such that 9 contiguous circle pixels are all darker than v - m or all brighter than v + m
      int best = threshold;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        int mn = pv[k], mx = pv[k];
#pragma unroll
        for (int j = 1; j < 9; j++) {
          mn = min(mn, pv[(k + j) & 15]);
          mx = max(mx, pv[(k + j) & 15]);
        }
        best = max(best, max(v - mx, mn - v));
      }
      sc = best - 1;
    }
  }
  score[(size_t)y * w + x] = sc;
}

__global__ void fast_nms_kernel(const int* __restrict__ score, int w, int h, int nonmax, uint8_t* __restrict__ flag) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const size_t i = (size_t)y * w + x;
  const int s = score[i];
  bool keep = s > 0;
  if (keep && nonmax) {
    // corners live in [3, w-3) x [3, h-3): the 8 neighbours exist
    keep = s > score[i - 1] && s > score[i + 1] && s > score[i - w - 1] && s > score[i - w] && s > score[i - w + 1] &&
           s > score[i + w - 1] && s > score[i + w] && s > score[i + w + 1];
  }
  flag[i] = keep ? 1 : 0;
}

__global__ void fast_gather_kernel(const int* __restrict__ sel, const int* __restrict__ d_n, int cap, const int* __restrict__ score,
                                   int w, float* __restrict__ xy, float* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(*d_n, cap)) return;
  const int q = sel[i];
  xy[2 * i] = (float)(q % w);
  xy[2 * i + 1] = (float)(q / w);
  sc[i] = (float)score[q];
}

// orb.cpp: half-width of row v of the circular patch of radius 15
__constant__ int c_orb_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2(y, x), scalar float path
__device__ __forceinline__ float orb_fast_atan2(float y, float x) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
  const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
  const float ax = fabsf(x), ay = fabsf(y), eps = 2.220446049250313e-16f;
  const bool xs = ax >= ay;
  const float c = xs ? __fdiv_rn(ay, __fadd_rn(ax, eps)) : __fdiv_rn(ax, __fadd_rn(ay, eps));
  const float c2 = __fmul_rn(c, c);
  float a = __fadd_rn(__fmul_rn(p7, c2), p5);
  a = __fadd_rn(__fmul_rn(a, c2), p3);
  a = __fadd_rn(__fmul_rn(a, c2), p1);
  a = __fmul_rn(a, c);
  if (!xs) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// ICAngles (orb.cpp): one warp per keypoint, lane l = column u = l - 15 (lane 31 idles)
__global__ void orb_angle_kernel(const uint8_t* __restrict__ img, int w, int h, const float* __restrict__ xy, int n,
                                 float* __restrict__ ang) {
  const int kp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (kp >= n) return;
  const int cx = __float2int_rn(xy[2 * kp]), cy = __float2int_rn(xy[2 * kp + 1]);
  const int u = lane - 15;
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const uint8_t* c = img + (size_t)cy * w + cx + u;
    int col = c[0];                         // sum of the column (for m_10)
    const int au = abs(u);
    for (int v = 1; v <= 15; v++) {
      if (au <= c_orb_umax[v]) {
        const int vp = c[v * w], vm = c[-v * w];
        col += vp + vm;
        m01 += v * (vp - vm);
      }
    }
    m10 = u * col;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  if (lane == 0) ang[kp] = orb_fast_atan2((float)m01, (float)m10);
}

// cv::resize(INTER_LINEAR_EXACT), 8-bit: horizontal pass in 8.8 fixed point, vertical pass in 16.16, rounded half up;
// offsets and weights come from the host (they are computed in double, orb_exact_coeffs)
__global__ void orb_resize_exact_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh,
                                        const int* __restrict__ ox, const int* __restrict__ cx, const int* __restrict__ oy,
                                        const int* __restrict__ cy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= dw) return;
  const int x0 = ox[x], x1 = min(x0 + 1, sw - 1), wx = cx[x];
  const int y0 = oy[y], y1 = min(y0 + 1, sh - 1), wy = cy[y];
  const uint8_t* r0 = src + (size_t)y0 * sw;
  const uint8_t* r1 = src + (size_t)y1 * sw;
  const int h0 = (int)r0[x0] * (256 - wx) + (int)r0[x1] * wx;
  const int h1 = (int)r1[x0] * (256 - wx) + (int)r1[x1] * wx;
  dst[(size_t)y * dw + x] = (uint8_t)((h0 * (256 - wy) + h1 * wy + 32768) >> 16);
}

// HarrisResponses (orb.cpp): the response ORB ranks its keypoints by.  A warp per keypoint, lanes over the 49 window
// pixels, integer sums reduced by shuffles, the float formula in OpenCV's order.
__global__ void orb_harris_kernel(const uint8_t* __restrict__ img, int w, int h, const float* __restrict__ xy, int n,
                                  float harris_k, float* __restrict__ resp) {
  const int kp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (kp >= n) return;
  const int x0 = __float2int_rn(xy[2 * kp]), y0 = __float2int_rn(xy[2 * kp + 1]);
  long long a = 0, b = 0, c = 0;
  for (int q = lane; q < 49; q += 32) {
    const uint8_t* p = img + (size_t)(y0 - 3 + q / 7) * w + (x0 - 3 + q % 7);
    const int ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-w + 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[w - 1]);
    const int iy = ((int)p[w] - (int)p[-w]) * 2 + ((int)p[w - 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[-w + 1]);
    a += ix * ix;
    b += iy * iy;
    c += ix * iy;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) {
    const float scale = __fdiv_rn(1.f, __fmul_rn(28.f, 255.f));       // 1 / ((1 << 2) * blockSize * 255)
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float af = (float)(int)a, bf = (float)(int)b, cf = (float)(int)c;   // OpenCV accumulates in int
    float t = __fsub_rn(__fmul_rn(af, bf), __fmul_rn(cf, cf));
    const float apb = __fadd_rn(af, bf);
    t = __fsub_rn(t, __fmul_rn(__fmul_rn(harris_k, apb), apb));
    resp[kp] = __fmul_rn(t, s4);
  }
}

// computeOrbDescriptors (orb.cpp), WTA_K = 2
__global__ void orb_describe_kernel(const uint8_t* __restrict__ sm, int w, int h, const float* __restrict__ xy,
                                    const float* __restrict__ ang, int n, uint8_t* __restrict__ desc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int kp = t >> 5, byte = t & 31;
  if (kp >= n) return;
  float angle = ang[kp];
  angle = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.0));      // angle *= (float)(CV_PI / 180.f)
  const float a = (float)cos((double)angle), b = (float)sin((double)angle);
  const int cx = __float2int_rn(xy[2 * kp]), cy = __float2int_rn(xy[2 * kp + 1]);
  const uint8_t* center = sm + (size_t)cy * w + cx;
  unsigned val = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const signed char* p = c_orb_pattern[8 * byte + k];
    int v[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float px = (float)p[2 * e], py = (float)p[2 * e + 1];
      const float x = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
      const float y = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
      v[e] = center[__float2int_rn(y) * w + __float2int_rn(x)];
    }
    val |= (unsigned)(v[0] < v[1]) << k;
  }
  desc[(size_t)kp * 32 + byte] = (uint8_t)val;
}


// ======================================================================================== all levels at once
// vo_orb_detect_and_compute keeps every selection of ORB::detectAndCompute on the device: one launch per stage covers
// all pyramid levels (blockIdx.z = level, or a flat index decoded through the per-level counts), and the host
// synchronises once, when the packed result has arrived.
constexpr int ORB_NL = 8, ORB_EDGE = 31, ORB_FAST_T = 20;

struct OrbLv {
  const uint8_t* img[ORB_NL];
  int w[ORB_NL], h[ORB_NL];
  int px_ofs[ORB_NL + 1];      // level offsets inside the concatenated per-pixel buffers
  int cand_ofs[ORB_NL + 1];    // level slices of the candidate buffers
  int quota[ORB_NL];
  int active[ORB_NL];
  float scale[ORB_NL];
};

__device__ __forceinline__ int fast_score_px(const uint8_t* __restrict__ img, int w, int h, int x, int y, int threshold) {
  int sc = 0;
  if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
    const uint8_t* p = img + (size_t)y * w + x;
    const int v = p[0];
    int pv[16];
    unsigned dark = 0u, bright = 0u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      pv[k] = p[c_fast_dy[k] * w + c_fast_dx[k]];
      dark |= (unsigned)(pv[k] < v - threshold) << k;
      bright |= (unsigned)(pv[k] > v + threshold) << k;
    }
    if (fast_has9(dark) || fast_has9(bright)) {
      int best = threshold;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        int mn = pv[k], mx = pv[k];
#pragma unroll
        for (int j = 1; j < 9; j++) {
          mn = min(mn, pv[(k + j) & 15]);
          mx = max(mx, pv[(k + j) & 15]);
        }
        best = max(best, max(v - mx, mn - v));
      }
      sc = best - 1;
    }
  }
  return sc;
}

__global__ void orb_fast_score_all_kernel(const OrbLv lv, int* __restrict__ score) {
  const int l = blockIdx.z, x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int w = lv.w[l], h = lv.h[l];
  if (x >= w || y >= h) return;
  score[lv.px_ofs[l] + y * w + x] = lv.active[l] ? fast_score_px(lv.img[l], w, h, x, y, ORB_FAST_T) : 0;
}

__global__ void orb_fast_nms_all_kernel(const OrbLv lv, const int* __restrict__ score_all, uint8_t* __restrict__ flag) {
  const int l = blockIdx.z, x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int w = lv.w[l], h = lv.h[l];
  if (x >= w || y >= h) return;
  const int* score = score_all + lv.px_ofs[l];
  const int i = y * w + x;
  const int s = score[i];
  bool keep = s > 0;
  if (keep)   // corners live in [3, w-3) x [3, h-3): the 8 neighbours exist
    keep = s > score[i - 1] && s > score[i + 1] && s > score[i - w - 1] && s > score[i - w] && s > score[i - w + 1] &&
           s > score[i + w - 1] && s > score[i + w] && s > score[i + w + 1];
  flag[lv.px_ofs[l] + i] = keep ? 1 : 0;
}

constexpr int ORB_SEL_T = 1024;

// Ordered append of one tile (one element per thread) by a whole CTA of ORB_SEL_T threads: returns the output position
// of the calling thread's element (valid if keep) and advances *base (shared) by the tile's count.
__device__ __forceinline__ int orb_tile_append(bool keep, int* s_wcnt /*32*/, int* s_base) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_wcnt[w] = __popc(b);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll 8
  for (int k = 0; k < ORB_SEL_T / 32; k++) {
    const int cnt = s_wcnt[k];
    before += k < w ? cnt : 0;
    total += cnt;
  }
  const int pos = *s_base + before + __popc(b & ((1u << lane) - 1u));
  __syncthreads();
  if (t == 0) *s_base += total;
  __syncthreads();
  return pos;
}

// first position in the ascending list sel[0..n) whose value is >= v
__device__ __forceinline__ int orb_lower_bound(const int* __restrict__ sel, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sel[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Per level (one CTA): cv::FAST's corner list of the level (its range of the raster-ordered selection) ->
// KeyPointsFilter::runByImageBorder(edgeThreshold) -> retainBest(2 * quota) by FAST score, ties kept: the scores are
// small integers, so the n-th largest is read off a 256-bin histogram.  Survivors keep their raster order.
__global__ void __launch_bounds__(ORB_SEL_T)
orb_select1_kernel(const OrbLv lv, const int* __restrict__ sel, const int* __restrict__ n_sel, const int* __restrict__ score,
                   float2* __restrict__ c_xy, int* __restrict__ hdr) {
  __shared__ int s_hist[256];
  __shared__ int s_wcnt[32];
  __shared__ int s_base, s_thr, s_lo, s_hi;
  const int l = blockIdx.x, t = threadIdx.x;
  const int w = lv.w[l], h = lv.h[l];
  if (t < 256) s_hist[t] = 0;
  if (t == 0) {
    const int n = *n_sel;
    s_lo = orb_lower_bound(sel, n, lv.px_ofs[l]);
    s_hi = orb_lower_bound(sel, n, lv.px_ofs[l + 1]);
    s_base = 0;
    if (l == 0) hdr[16] = n;
  }
  __syncthreads();
  const int lo = s_lo, hi = lv.active[l] ? s_hi : s_lo;
  for (int i = lo + t; i < hi; i += ORB_SEL_T) {
    const int q = sel[i] - lv.px_ofs[l];
    const int x = q % w, y = q / w;
    if (x >= ORB_EDGE && x < w - ORB_EDGE && y >= ORB_EDGE && y < h - ORB_EDGE) atomicAdd(&s_hist[min(score[lv.px_ofs[l] + q], 255)], 1);
  }
  __syncthreads();
  if (t == 0) {
    const int target = 2 * lv.quota[l];
    int m = 0;
    for (int k = 0; k < 256; k++) m += s_hist[k];
    int thr = 0;
    if (m > target) {
      int cum = 0;
      for (thr = 255; thr > 0; thr--) {
        cum += s_hist[thr];
        if (cum >= target) break;
      }
    }
    s_thr = target > 0 ? thr : 256;
  }
  __syncthreads();
  const int thr = s_thr;
  float2* out = c_xy + lv.cand_ofs[l];
  for (int i0 = lo; i0 < hi; i0 += ORB_SEL_T) {
    const int i = i0 + t;
    bool keep = false;
    int x = 0, y = 0;
    if (i < hi) {
      const int q = sel[i] - lv.px_ofs[l];
      x = q % w;
      y = q / w;
      keep = x >= ORB_EDGE && x < w - ORB_EDGE && y >= ORB_EDGE && y < h - ORB_EDGE && min(score[lv.px_ofs[l] + q], 255) >= thr;
    }
    const int pos = orb_tile_append(keep, s_wcnt, &s_base);
    if (keep) out[pos] = make_float2((float)x, (float)y);
  }
  if (t == 0) hdr[8 + l] = s_base;
}

// flat index over the per-level counts n[0..8) -> (level, index inside the level); false past the end
__device__ __forceinline__ bool orb_decode(const int* __restrict__ n, int k, int& l, int& i, int& before) {
  before = 0;
#pragma unroll
  for (int j = 0; j < ORB_NL; j++) {
    const int c = n[j];
    if (k < before + c) {
      l = j;
      i = k - before;
      return true;
    }
    before += c;
  }
  return false;
}

__device__ __forceinline__ float orb_harris_warp(const uint8_t* __restrict__ img, int w, int x0, int y0, float harris_k, int lane) {
  long long a = 0, b = 0, c = 0;
  for (int q = lane; q < 49; q += 32) {
    const uint8_t* p = img + (size_t)(y0 - 3 + q / 7) * w + (x0 - 3 + q % 7);
    const int ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-w + 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[w - 1]);
    const int iy = ((int)p[w] - (int)p[-w]) * 2 + ((int)p[w - 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[-w + 1]);
    a += ix * ix;
    b += iy * iy;
    c += ix * iy;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const float scale = __fdiv_rn(1.f, __fmul_rn(28.f, 255.f));       // 1 / ((1 << 2) * blockSize * 255)
  const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
  const float af = (float)(int)a, bf = (float)(int)b, cf = (float)(int)c;   // OpenCV accumulates in int
  float t = __fsub_rn(__fmul_rn(af, bf), __fmul_rn(cf, cf));
  const float apb = __fadd_rn(af, bf);
  t = __fsub_rn(t, __fmul_rn(__fmul_rn(harris_k, apb), apb));
  return __fmul_rn(t, s4);
}

// HarrisResponses of the first selection's survivors, all levels: a warp per keypoint, grid-stride
__global__ void orb_harris_all_kernel(const OrbLv lv, const int* __restrict__ hdr, const float2* __restrict__ c_xy,
                                      float* __restrict__ c_resp) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;; k += nwarps) {
    int l, i, before;
    if (!orb_decode(hdr + 8, k, l, i, before)) break;
    const float2 p = c_xy[lv.cand_ofs[l] + i];
    const float r = orb_harris_warp(lv.img[l], lv.w[l], __float2int_rn(p.x), __float2int_rn(p.y), 0.04f, lane);
    if (lane == 0) c_resp[lv.cand_ofs[l] + i] = r;
  }
}

// order-preserving key of a float (-0 counts as +0, like the float comparison it stands for)
__device__ __forceinline__ unsigned orb_float_key(float f) {
  unsigned u = __float_as_uint(f == 0.f ? 0.f : f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Per level (one CTA): retainBest(quota) by Harris response -- the quota-th largest response by an exact radix select
// on the float keys (four 8-bit passes), everything >= it stays (ties included), raster order kept, in place.
__global__ void __launch_bounds__(ORB_SEL_T)
orb_select2_kernel(const OrbLv lv, float2* __restrict__ c_xy, float* __restrict__ c_resp, int* __restrict__ hdr) {
  __shared__ int s_hist[256];
  __shared__ int s_wcnt[32];
  __shared__ int s_base, s_k;
  __shared__ unsigned s_prefix;
  const int l = blockIdx.x, t = threadIdx.x;
  const int n = hdr[8 + l], quota = lv.quota[l];
  float2* xy = c_xy + lv.cand_ofs[l];
  float* resp = c_resp + lv.cand_ofs[l];
  unsigned thr_key = 0;     // keep everything
  if (n > quota && quota > 0) {
    if (t == 0) {
      s_prefix = 0;
      s_k = quota;
    }
    for (int pass = 0; pass < 4; pass++) {
      const int shift = 24 - 8 * pass;
      if (t < 256) s_hist[t] = 0;
      __syncthreads();
      const unsigned prefix = s_prefix, pmask = pass == 0 ? 0u : 0xffffffffu << (shift + 8);
      for (int i = t; i < n; i += ORB_SEL_T) {
        const unsigned key = orb_float_key(resp[i]);
        if ((key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      if (t == 0) {
        int k = s_k, d = 255;
        for (; d > 0; d--) {
          if (s_hist[d] >= k) break;
          k -= s_hist[d];
        }
        s_k = k;
        s_prefix = prefix | ((unsigned)d << shift);
      }
      __syncthreads();
    }
    thr_key = s_prefix;
  }
  if (t == 0) s_base = 0;
  __syncthreads();
  const int n_in = quota > 0 ? n : 0;
  for (int i0 = 0; i0 < n_in; i0 += ORB_SEL_T) {
    const int i = i0 + t;
    bool keep = false;
    float2 p = make_float2(0.f, 0.f);
    float r = 0.f;
    if (i < n_in) {
      p = xy[i];
      r = resp[i];
      keep = orb_float_key(r) >= thr_key;
    }
    const int pos = orb_tile_append(keep, s_wcnt, &s_base);   // (its barriers separate this tile's reads from its writes)
    if (keep) {
      xy[pos] = p;
      resp[pos] = r;
    }
  }
  if (t == 0) hdr[l] = s_base;
}

// ICAngles of the final keypoints of all levels (warp per keypoint, grid-stride); lane 0 also writes the keypoint's
// packed position (level coordinates times the level scale) and response
__global__ void orb_angle_all_kernel(const OrbLv lv, const int* __restrict__ hdr, const float2* __restrict__ c_xy,
                                     const float* __restrict__ c_resp, int out_cap, float2* __restrict__ o_xy,
                                     float* __restrict__ o_resp, float* __restrict__ o_ang) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < out_cap; k += nwarps) {
    int l, i, before;
    if (!orb_decode(hdr, k, l, i, before)) break;
    const float2 p = c_xy[lv.cand_ofs[l] + i];
    const int w = lv.w[l];
    const int cx = __float2int_rn(p.x), cy = __float2int_rn(p.y);
    const int u = lane - 15;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
      const uint8_t* c = lv.img[l] + (size_t)cy * w + cx + u;
      int col = c[0];
      const int au = abs(u);
      for (int v = 1; v <= 15; v++) {
        if (au <= c_orb_umax[v]) {
          const int vp = c[v * w], vm = c[-v * w];
          col += vp + vm;
          m01 += v * (vp - vm);
        }
      }
      m10 = u * col;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    if (lane == 0) {
      o_ang[k] = orb_fast_atan2((float)m01, (float)m10);
      o_xy[k] = make_float2(__fmul_rn(p.x, lv.scale[l]), __fmul_rn(p.y, lv.scale[l]));   // allKeypoints[i].pt *= scale
      o_resp[k] = c_resp[lv.cand_ofs[l] + i];
    }
  }
}

__global__ void orb_smooth_rows_all_kernel(const OrbLv lv, float* __restrict__ rowf) {
  const int l = blockIdx.z, x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int w = lv.w[l], h = lv.h[l];
  if (x >= w || y >= h || !lv.active[l]) return;
  const uint8_t* r = lv.img[l] + (size_t)y * w;
  float s = __fmul_rn(c_orb_gauss[0], (float)r[orb_reflect101(x - 3, w)]);
#pragma unroll
  for (int j = 1; j < 7; j++) s = __fmaf_rn(c_orb_gauss[j < 4 ? j : 6 - j], (float)r[orb_reflect101(x - 3 + j, w)], s);
  rowf[lv.px_ofs[l] + y * w + x] = s;
}

__global__ void orb_smooth_cols_all_kernel(const OrbLv lv, const float* __restrict__ rowf_all, uint8_t* __restrict__ out) {
  const int l = blockIdx.z, x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int w = lv.w[l], h = lv.h[l];
  if (x >= w || y >= h || !lv.active[l]) return;
  const float* rowf = rowf_all + lv.px_ofs[l];
  float t = __fmul_rn(c_orb_gauss[3], rowf[(size_t)y * w + x]);
#pragma unroll
  for (int j = 1; j < 4; j++) {
    const float a = rowf[(size_t)orb_reflect101(y + j, h) * w + x], b = rowf[(size_t)orb_reflect101(y - j, h) * w + x];
    t = __fmaf_rn(c_orb_gauss[3 - j], __fadd_rn(a, b), t);
  }
  const int v = __float2int_rn(t);
  out[lv.px_ofs[l] + y * w + x] = (uint8_t)min(max(v, 0), 255);
}

// computeOrbDescriptors of the final keypoints of all levels: a thread per (keypoint, descriptor byte), grid-stride
__global__ void orb_describe_all_kernel(const OrbLv lv, const int* __restrict__ hdr, const float2* __restrict__ c_xy,
                                        const float* __restrict__ o_ang, const uint8_t* __restrict__ sm_all, int out_cap,
                                        uint8_t* __restrict__ o_desc) {
  const int byte = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < out_cap; k += nwarps) {
    int l, i, before;
    if (!orb_decode(hdr, k, l, i, before)) break;
    const float2 p = c_xy[lv.cand_ofs[l] + i];
    const int w = lv.w[l];
    float angle = o_ang[k];
    angle = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.0));      // angle *= (float)(CV_PI / 180.f)
    const float a = (float)cos((double)angle), b = (float)sin((double)angle);
    const uint8_t* center = sm_all + lv.px_ofs[l] + (size_t)__float2int_rn(p.y) * w + __float2int_rn(p.x);
    unsigned val = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const signed char* pt = c_orb_pattern[8 * byte + q];
      int v[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float px = (float)pt[2 * e], py = (float)pt[2 * e + 1];
        const float x = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
        const float y = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
        v[e] = center[__float2int_rn(y) * w + __float2int_rn(x)];
      }
      val |= (unsigned)(v[0] < v[1]) << q;
    }
    o_desc[(size_t)k * 32 + byte] = (uint8_t)val;
  }
}

static int orb_ensure(vo_ctx* c, int w, int h, int n) {
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  if (!o) {
    o = new Orb();
    c->orb = o;
    VO_CUDA(cudaMemcpyToSymbol(c_orb_pattern, ORB_PATTERN_31, sizeof(ORB_PATTERN_31)));
    VO_CUDA(cudaMemcpyToSymbol(c_orb_gauss, ORB_GAUSS_7_2, sizeof(ORB_GAUSS_7_2)));
  }
  if ((size_t)w * h > (size_t)o->w * o->h) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->img);
    cudaFree(o->rowf);
    cudaFree(o->sm);
    o->img = nullptr; o->rowf = nullptr; o->sm = nullptr;
    const size_t npx = (size_t)w * h;
    VO_CUDA(cudaMalloc(&o->img, npx));
    VO_CUDA(cudaMalloc(&o->rowf, npx * sizeof(float)));
    VO_CUDA(cudaMalloc(&o->sm, npx));
    o->w = w;
    o->h = h;
  }
  if (n > o->cap) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->xy);
    cudaFree(o->ang);
    cudaFree(o->desc);
    o->xy = nullptr; o->ang = nullptr; o->desc = nullptr;
    VO_CUDA(cudaMalloc(&o->xy, (size_t)n * 2 * sizeof(float)));
    VO_CUDA(cudaMalloc(&o->ang, (size_t)n * sizeof(float)));
    VO_CUDA(cudaMalloc(&o->desc, (size_t)n * 32));
    o->cap = n;
  }
  return VO_OK;
}

// the per-pixel buffers of the FAST stage follow the image buffers' size
static int orb_ensure_fast(vo_ctx* c, Orb* o) {
  if (!o->d_n) VO_CUDA(cudaMalloc(&o->d_n, 16 * sizeof(int)));
  const size_t npx = (size_t)o->w * o->h;
  if (o->px_cap >= npx) return VO_OK;
  VO_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(o->score); cudaFree(o->flag); cudaFree(o->sel); cudaFree(o->cub_tmp);
  o->score = nullptr; o->flag = nullptr; o->sel = nullptr; o->cub_tmp = nullptr;
  o->px_cap = 0;
  VO_CUDA(cudaMalloc(&o->score, npx * sizeof(int)));
  VO_CUDA(cudaMalloc(&o->flag, npx));
  VO_CUDA(cudaMalloc(&o->sel, npx * sizeof(int)));
  size_t tb = 0;
  VO_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, cub::CountingInputIterator<int>(0), (const uint8_t*)nullptr, (int*)nullptr,
                                     (int*)nullptr, (int)npx, c->stream));
  o->cub_bytes = tb;
  VO_CUDA(cudaMalloc(&o->cub_tmp, tb + 256));
  o->px_cap = npx;
  return VO_OK;
}

// buffers of the all-level pipeline (vo_orb_detect_and_compute)
static int orb_ensure_all(vo_ctx* c, Orb* o, size_t all_px, int n_cand, int out_cap) {
  if (!o->d_hdr) {
    VO_CUDA(cudaMalloc(&o->d_hdr, 32 * sizeof(int)));
    VO_CUDA(cudaMemsetAsync(o->d_hdr, 0, 32 * sizeof(int), c->stream));
  }
  if (all_px > o->all_px) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->a_score); cudaFree(o->a_flag); cudaFree(o->a_sel); cudaFree(o->a_rowf); cudaFree(o->a_sm); cudaFree(o->a_cub);
    o->a_score = nullptr; o->a_flag = nullptr; o->a_sel = nullptr; o->a_rowf = nullptr; o->a_sm = nullptr; o->a_cub = nullptr;
    o->all_px = 0;
    VO_CUDA(cudaMalloc(&o->a_score, all_px * sizeof(int)));
    VO_CUDA(cudaMalloc(&o->a_flag, all_px));
    VO_CUDA(cudaMalloc(&o->a_sel, all_px * sizeof(int)));
    VO_CUDA(cudaMalloc(&o->a_rowf, all_px * sizeof(float)));
    VO_CUDA(cudaMalloc(&o->a_sm, all_px));
    size_t tb = 0;
    VO_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, cub::CountingInputIterator<int>(0), (const uint8_t*)nullptr, (int*)nullptr,
                                       (int*)nullptr, (int)all_px, c->stream));
    o->a_cub_bytes = tb;
    VO_CUDA(cudaMalloc(&o->a_cub, tb + 256));
    o->all_px = all_px;
  }
  if (n_cand > o->n_cand) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->c_xy); cudaFree(o->c_resp);
    o->c_xy = nullptr; o->c_resp = nullptr;
    o->n_cand = 0;
    VO_CUDA(cudaMalloc(&o->c_xy, (size_t)n_cand * sizeof(float2)));
    VO_CUDA(cudaMalloc(&o->c_resp, (size_t)n_cand * sizeof(float)));
    o->n_cand = n_cand;
  }
  if (out_cap > o->out_cap) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->o_xy); cudaFree(o->o_resp); cudaFree(o->o_ang); cudaFree(o->o_desc);
    o->o_xy = nullptr; o->o_resp = nullptr; o->o_ang = nullptr; o->o_desc = nullptr;
    o->out_cap = 0;
    VO_CUDA(cudaMalloc(&o->o_xy, (size_t)out_cap * sizeof(float2)));
    VO_CUDA(cudaMalloc(&o->o_resp, (size_t)out_cap * sizeof(float)));
    VO_CUDA(cudaMalloc(&o->o_ang, (size_t)out_cap * sizeof(float)));
    VO_CUDA(cudaMalloc(&o->o_desc, (size_t)out_cap * 32));
    o->out_cap = out_cap;
  }
  return VO_OK;
}


static int orb_smooth_enqueue(vo_ctx* c, Orb* o, const uint8_t* img, int stride, int w, int h) {
  VO_CUDA(cudaMemcpy2DAsync(o->img, w, img, stride, w, h, cudaMemcpyDefault, c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_smooth_rows_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(o->img, w, h, o->rowf);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_smooth_cols_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(o->rowf, w, h, o->sm);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo

using namespace vo;

int vo_orb_smooth(vo_ctx* c, const uint8_t* img, int stride, int width, int height, uint8_t* out, int out_stride) {
  if (!c) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaSetDevice(c->device));
  if (!img || !out || width < 7 || height < 7 || stride < width || out_stride < width) return VO_ERR_INVALID_ARG;
  VO_TRY(orb_ensure(c, width, height, 0));
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  VO_TRY(orb_smooth_enqueue(c, o, img, stride, width, height));
  VO_CUDA(cudaMemcpy2DAsync(out, out_stride, o->sm, width, width, height, cudaMemcpyDefault, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

int vo_orb_describe(vo_ctx* c, const uint8_t* img, int stride, int width, int height, const float* xy,
                    const float* angle_deg, int n, uint8_t* desc) {
  if (!c) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaSetDevice(c->device));
  if (!img || width < 64 || height < 64 || stride < width || n < 0 || (n > 0 && (!xy || !desc))) return VO_ERR_INVALID_ARG;
  if (n == 0) return VO_OK;
  // a rotated test point reaches 19 px from the rounded centre; cv2 itself drops keypoints closer than 31 px
  for (int i = 0; i < n; i++) {
    const float x = xy[2 * i], y = xy[2 * i + 1];
    if (!(x >= 19.5f && x <= (float)width - 20.5f && y >= 19.5f && y <= (float)height - 20.5f)) {
      set_error("vo_orb_describe: keypoint %d (%.2f, %.2f) is closer than 19.5 px to the image border", i, x, y);
      return VO_ERR_INVALID_ARG;
    }
  }
  VO_TRY(orb_ensure(c, width, height, n));
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  VO_TRY(orb_smooth_enqueue(c, o, img, stride, width, height));
  VO_CUDA(cudaMemcpyAsync(o->xy, xy, (size_t)n * 2 * sizeof(float), cudaMemcpyDefault, c->stream));
  if (angle_deg) {
    VO_CUDA(cudaMemcpyAsync(o->ang, angle_deg, (size_t)n * sizeof(float), cudaMemcpyDefault, c->stream));
  } else {   // what detectAndCompute does: IC_Angle on the unsmoothed level
    LaunchScope ls(c, VO_K_MISC);
    orb_angle_kernel<<<div_up(n * 32, 256), 256, 0, c->stream>>>(o->img, width, height, o->xy, n, o->ang);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_describe_kernel<<<div_up(n * 32, 256), 256, 0, c->stream>>>(o->sm, width, height, o->xy, o->ang, n, o->desc);
  }
  VO_CUDA(cudaGetLastError());
  VO_CUDA(cudaMemcpyAsync(desc, o->desc, (size_t)n * 32, cudaMemcpyDefault, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

int vo_orb_angles(vo_ctx* c, const uint8_t* img, int stride, int width, int height, const float* xy, int n,
                  float* angle_deg) {
  if (!c) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaSetDevice(c->device));
  if (!img || width < 64 || height < 64 || stride < width || n < 0 || (n > 0 && (!xy || !angle_deg))) return VO_ERR_INVALID_ARG;
  if (n == 0) return VO_OK;
  for (int i = 0; i < n; i++) {
    const float x = xy[2 * i], y = xy[2 * i + 1];
    if (!(x >= 15.5f && x <= (float)width - 16.5f && y >= 15.5f && y <= (float)height - 16.5f)) {
      set_error("vo_orb_angles: keypoint %d (%.2f, %.2f) is closer than 15.5 px to the image border", i, x, y);
      return VO_ERR_INVALID_ARG;
    }
  }
  VO_TRY(orb_ensure(c, width, height, n));
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  VO_CUDA(cudaMemcpy2DAsync(o->img, width, img, stride, width, height, cudaMemcpyDefault, c->stream));
  VO_CUDA(cudaMemcpyAsync(o->xy, xy, (size_t)n * 2 * sizeof(float), cudaMemcpyDefault, c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_angle_kernel<<<div_up(n * 32, 256), 256, 0, c->stream>>>(o->img, width, height, o->xy, n, o->ang);
  }
  VO_CUDA(cudaGetLastError());
  VO_CUDA(cudaMemcpyAsync(angle_deg, o->ang, (size_t)n * sizeof(float), cudaMemcpyDefault, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

int vo_orb_harris(vo_ctx* c, const uint8_t* img, int stride, int width, int height, const float* xy, int n, float* response) {
  if (!c) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaSetDevice(c->device));
  if (!img || width < 64 || height < 64 || stride < width || n < 0 || (n > 0 && (!xy || !response))) return VO_ERR_INVALID_ARG;
  if (n == 0) return VO_OK;
  for (int i = 0; i < n; i++) {
    const float x = xy[2 * i], y = xy[2 * i + 1];
    if (!(x >= 4.5f && x <= (float)width - 5.5f && y >= 4.5f && y <= (float)height - 5.5f)) {
      set_error("vo_orb_harris: keypoint %d (%.2f, %.2f) is closer than 4.5 px to the image border", i, x, y);
      return VO_ERR_INVALID_ARG;
    }
  }
  VO_TRY(orb_ensure(c, width, height, n));
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  VO_CUDA(cudaMemcpy2DAsync(o->img, width, img, stride, width, height, cudaMemcpyDefault, c->stream));
  VO_CUDA(cudaMemcpyAsync(o->xy, xy, (size_t)n * 2 * sizeof(float), cudaMemcpyDefault, c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_harris_kernel<<<div_up(n * 32, 256), 256, 0, c->stream>>>(o->img, width, height, o->xy, n, 0.04f, o->ang);
  }
  VO_CUDA(cudaGetLastError());
  VO_CUDA(cudaMemcpyAsync(response, o->ang, (size_t)n * sizeof(float), cudaMemcpyDefault, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

int vo_fast9(vo_ctx* c, const uint8_t* img, int stride, int width, int height, int threshold, int nonmax_suppression,
             float* xy, float* score, int cap, int* n_out) {
  if (!c) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaSetDevice(c->device));
  if (!img || !n_out || width < 7 || height < 7 || stride < width || threshold < 1 || threshold > 254 || cap < 0 ||
      (cap > 0 && !xy))
    return VO_ERR_INVALID_ARG;
  *n_out = 0;
  VO_TRY(orb_ensure(c, width, height, std::max(cap, 1)));
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  const int n = width * height;
  VO_TRY(orb_ensure_fast(c, o));
  VO_CUDA(cudaMemcpy2DAsync(o->img, width, img, stride, width, height, cudaMemcpyDefault, c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    fast_score_kernel<<<dim3(div_up(width, 128), height), 128, 0, c->stream>>>(o->img, width, height, threshold, o->score);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    fast_nms_kernel<<<dim3(div_up(width, 128), height), 128, 0, c->stream>>>(o->score, width, height, nonmax_suppression, o->flag);
  }
  size_t tb = o->cub_bytes;
  VO_CUDA(cub::DeviceSelect::Flagged(o->cub_tmp, tb, cub::CountingInputIterator<int>(0), o->flag, o->sel, o->d_n, n, c->stream));
  c->launch_count++;
  if (cap > 0) {
    LaunchScope ls(c, VO_K_MISC);
    fast_gather_kernel<<<div_up(cap, 256), 256, 0, c->stream>>>(o->sel, o->d_n, cap, o->score, width, o->xy, o->ang);
  }
  VO_CUDA(cudaGetLastError());
  int hn = 0;
  VO_CUDA(cudaMemcpyAsync(&hn, o->d_n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  *n_out = hn;
  const int k = std::min(hn, cap);
  if (k > 0) {
    VO_CUDA(cudaMemcpyAsync(xy, o->xy, (size_t)k * 2 * sizeof(float), cudaMemcpyDefault, c->stream));
    if (score) VO_CUDA(cudaMemcpyAsync(score, o->ang, (size_t)k * sizeof(float), cudaMemcpyDefault, c->stream));
    VO_CUDA(cudaStreamSynchronize(c->stream));
  }
  return hn > cap ? VO_ERR_CAPACITY : VO_OK;
}

// ---------------------------------------------------------------------------------------- ORB::detectAndCompute
namespace {

// resize.cpp interpolationLinear::getCoeffs (8-bit): IEEE double arithmetic, like OpenCV's softdouble
void orb_exact_coeffs(int srcsize, int dstsize, int* ofs, int* c1) {
  const double inv = (double)dstsize / (double)srcsize;
  const double scale = 1.0 / inv;
  for (int v = 0; v < dstsize; v++) {
    const double fval = scale * ((double)v + 0.5) - 0.5;
    const int ival = (int)std::floor(fval);
    ofs[v] = 0;
    c1[v] = 0;
    if (ival >= 0 && srcsize > 1) {
      if (ival < srcsize - 1) {
        ofs[v] = ival;
        c1[v] = (int)std::nearbyint((fval - (double)ival) * 256.0);
      } else {
        ofs[v] = srcsize - 1;
      }
    }
  }
}

}  // namespace

int vo_orb_detect_and_compute(vo_ctx* c, const uint8_t* img, int stride, int width, int height, int nfeatures, float* xy,
                              int32_t* octave, float* response, float* angle_deg, uint8_t* desc, int cap, int* n_out) {
  if (!c) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaSetDevice(c->device));
  if (!img || !n_out || width < 64 || height < 64 || stride < width || nfeatures < 1 || cap < 0 || (cap > 0 && (!xy || !desc)))
    return VO_ERR_INVALID_ARG;
  *n_out = 0;
  // ORB::create() defaults (the reference passes none, src/optimizationStuff.cpp:49): scaleFactor 1.2f kept in a
  // double, 8 levels, edgeThreshold 31, firstLevel 0, WTA_K 2, HARRIS_SCORE, patchSize 31, fastThreshold 20
  constexpr int NL = ORB_NL, EDGE = ORB_EDGE;
  const double scale_factor = (double)1.2f;
  float lscale[NL];
  int lw[NL], lh[NL], quota[NL];
  for (int l = 0; l < NL; l++) {
    lscale[l] = (float)std::pow(scale_factor, (double)l);
    lw[l] = (int)std::lrintf((float)width / lscale[l]);      // cvRound(image.cols / scale)
    lh[l] = (int)std::lrintf((float)height / lscale[l]);
  }
  {
    const float factor = (float)(1.0 / scale_factor);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)NL));
    int sum = 0;
    for (int l = 0; l < NL - 1; l++) {
      quota[l] = (int)std::lrintf(nd);
      sum += quota[l];
      nd *= factor;
    }
    quota[NL - 1] = std::max(nfeatures - sum, 0);
  }
  // per-level slices of the candidate buffers: after 3x3 suppression there is at most one corner per 2x2 pixels
  OrbLv lv;
  lv.px_ofs[0] = 0;
  lv.cand_ofs[0] = 0;
  for (int l = 0; l < NL; l++) {
    lv.w[l] = lw[l];
    lv.h[l] = lh[l];
    lv.quota[l] = quota[l];
    lv.scale[l] = lscale[l];
    lv.active[l] = lw[l] > 2 * EDGE && lh[l] > 2 * EDGE && quota[l] > 0;      // else the border filter leaves nothing
    lv.px_ofs[l + 1] = lv.px_ofs[l] + lw[l] * lh[l];
    lv.cand_ofs[l + 1] = lv.cand_ofs[l] + (lv.active[l] ? lw[l] * lh[l] / 4 + 16 : 0);
  }
  const int n_cand = std::max(lv.cand_ofs[NL], 1);
  const size_t all_px = (size_t)lv.px_ofs[NL];
  VO_TRY(orb_ensure(c, width, height, 1));
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  const int DESC_CAP = std::max(8192, 2 * nfeatures + 4096);     // keypoints the packed result holds
  VO_TRY(orb_ensure_all(c, o, all_px, n_cand, DESC_CAP));
  // pyramid levels 1.. and the resize tables
  size_t pyr_need = 0;
  for (int l = 1; l < NL; l++) pyr_need += (size_t)lw[l] * lh[l];
  if (pyr_need > o->pyr_bytes) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->pyr);
    o->pyr = nullptr;
    VO_CUDA(cudaMalloc(&o->pyr, pyr_need));
    o->pyr_bytes = pyr_need;
  }
  // the resize tables of all levels in one upload, then the whole pyramid: resize(prevImg, currImg, sz, 0, 0,
  // INTER_LINEAR_EXACT), each level from the previous one
  size_t coef_ofs[NL] = {0};
  size_t n_coef = 0;
  for (int l = 1; l < NL; l++) {
    coef_ofs[l] = n_coef;
    n_coef += 2 * (size_t)(lw[l] + lh[l]);
  }
  if ((int)n_coef > o->coef_cap) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->coef);
    o->coef = nullptr;
    VO_CUDA(cudaMalloc(&o->coef, n_coef * sizeof(int)));
    o->coef_cap = (int)n_coef;
  }
  // pinned scratch (pageable copies are staged by the driver at a few GB/s and serialise with the stream):
  // image | resize tables | header | packed results; every section 16-byte aligned
  const size_t pin_img = (((size_t)width * height) + 15) & ~(size_t)15, pin_coef = ((n_coef * sizeof(int)) + 15) & ~(size_t)15,
               pin_hdr = 32 * sizeof(int), pin_out = (size_t)DESC_CAP * (8 + 4 + 4 + 32);
  if (pin_img + pin_coef + pin_hdr + pin_out > o->h_pin_bytes) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFreeHost(o->h_pin);
    o->h_pin = nullptr;
    o->h_pin_bytes = 0;
    VO_CUDA(cudaMallocHost(&o->h_pin, pin_img + pin_coef + pin_hdr + pin_out));
    o->h_pin_bytes = pin_img + pin_coef + pin_hdr + pin_out;
  }
  int* p_coef = reinterpret_cast<int*>(o->h_pin + pin_img);
  int* p_hdr = reinterpret_cast<int*>(o->h_pin + pin_img + pin_coef);
  float* p_xy = reinterpret_cast<float*>(o->h_pin + pin_img + pin_coef + pin_hdr);
  float* p_resp = p_xy + 2 * (size_t)DESC_CAP;
  float* p_ang = p_resp + DESC_CAP;
  uint8_t* p_desc = reinterpret_cast<uint8_t*>(p_ang + DESC_CAP);
  for (int l = 1; l < NL; l++) {
    int* t = p_coef + coef_ofs[l];
    orb_exact_coeffs(lw[l - 1], lw[l], t, t + lw[l]);
    orb_exact_coeffs(lh[l - 1], lh[l], t + 2 * lw[l], t + 2 * lw[l] + lh[l]);
  }
  for (int y = 0; y < height; y++) memcpy(o->h_pin + (size_t)y * width, img + (size_t)y * stride, width);
  VO_CUDA(cudaMemcpyAsync(o->img, o->h_pin, (size_t)width * height, cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(o->coef, p_coef, n_coef * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  lv.img[0] = o->img;
  {
    uint8_t* next = o->pyr;
    for (int l = 1; l < NL; l++) {
      const int* t = o->coef + coef_ofs[l];
      LaunchScope ls(c, VO_K_MISC);
      orb_resize_exact_kernel<<<dim3(div_up(lw[l], 128), lh[l]), 128, 0, c->stream>>>(lv.img[l - 1], lw[l - 1], lh[l - 1], next, lw[l],
                                                                                    lh[l], t, t + lw[l], t + 2 * lw[l],
                                                                                    t + 2 * lw[l] + lh[l]);
      lv.img[l] = next;
      next += (size_t)lw[l] * lh[l];
    }
  }
  // ---- FAST (threshold 20, suppression) on every level: one score launch, one suppression launch, one raster-ordered
  // selection over the concatenated levels
  const dim3 g_px(div_up(width, 128), height, NL);
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_fast_score_all_kernel<<<g_px, 128, 0, c->stream>>>(lv, o->a_score);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_fast_nms_all_kernel<<<g_px, 128, 0, c->stream>>>(lv, o->a_score, o->a_flag);
  }
  {
    size_t tb = o->a_cub_bytes;
    VO_CUDA(cub::DeviceSelect::Flagged(o->a_cub, tb, cub::CountingInputIterator<int>(0), o->a_flag, o->a_sel, o->d_hdr + 17,
                                       (int)all_px, c->stream));
    c->launch_count++;
  }
  // ---- runByImageBorder + retainBest(2 * quota) by FAST score, HarrisResponses, retainBest(quota) by response
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_select1_kernel<<<NL, ORB_SEL_T, 0, c->stream>>>(lv, o->a_sel, o->d_hdr + 17, o->a_score, o->c_xy, o->d_hdr);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_harris_all_kernel<<<296, 256, 0, c->stream>>>(lv, o->d_hdr, o->c_xy, o->c_resp);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_select2_kernel<<<NL, ORB_SEL_T, 0, c->stream>>>(lv, o->c_xy, o->c_resp, o->d_hdr);
  }
  // ---- ICAngles on the unsmoothed levels, GaussianBlur of the levels, computeOrbDescriptors; the order inside a level is
  // raster (y, x), OpenCV's is what std::nth_element leaves
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_angle_all_kernel<<<296, 256, 0, c->stream>>>(lv, o->d_hdr, o->c_xy, o->c_resp, DESC_CAP, o->o_xy, o->o_resp, o->o_ang);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_smooth_rows_all_kernel<<<g_px, 128, 0, c->stream>>>(lv, o->a_rowf);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_smooth_cols_all_kernel<<<g_px, 128, 0, c->stream>>>(lv, o->a_rowf, o->a_sm);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    orb_describe_all_kernel<<<592, 256, 0, c->stream>>>(lv, o->d_hdr, o->c_xy, o->o_ang, o->a_sm, DESC_CAP, o->o_desc);
  }
  VO_CUDA(cudaGetLastError());
  // the packed result: the header and the first `guess` keypoints travel before the one synchronisation; a frame with
  // more (ties at a selection threshold) fetches the remainder afterwards
  const int guess = std::min(DESC_CAP, nfeatures + 512);
  auto fetch = [&](int from, int to) -> int {
    const size_t m = (size_t)(to - from);
    VO_CUDA(cudaMemcpyAsync(p_xy + 2 * (size_t)from, o->o_xy + from, m * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(p_resp + from, o->o_resp + from, m * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(p_ang + from, o->o_ang + from, m * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(p_desc + (size_t)from * 32, o->o_desc + (size_t)from * 32, m * 32, cudaMemcpyDeviceToHost, c->stream));
    return VO_OK;
  };
  VO_CUDA(cudaMemcpyAsync(p_hdr, o->d_hdr, 18 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VO_TRY(fetch(0, guess));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  int total = 0;
  for (int l = 0; l < NL; l++) total += p_hdr[l];
  if (total > DESC_CAP) {
    set_error("vo_orb_detect_and_compute: more than %d keypoints", DESC_CAP);
    return VO_ERR_CAPACITY;
  }
  if (total > guess) {
    VO_TRY(fetch(guess, total));
    VO_CUDA(cudaStreamSynchronize(c->stream));
  }
  const int n_copy = std::min(total, cap);
  if (n_copy > 0) {
    memcpy(xy, p_xy, (size_t)n_copy * 2 * sizeof(float));
    if (response) memcpy(response, p_resp, (size_t)n_copy * sizeof(float));
    if (angle_deg) memcpy(angle_deg, p_ang, (size_t)n_copy * sizeof(float));
    memcpy(desc, p_desc, (size_t)n_copy * 32);
    if (octave) {
      int k = 0;
      for (int l = 0; l < NL; l++)
        for (int i = 0; i < p_hdr[l] && k < n_copy; i++) octave[k++] = l;
    }
  }
  *n_out = total;
  return total > cap ? VO_ERR_CAPACITY : VO_OK;
}
