// orb.cu -- ORB for the loop detector (reference src/optimizationStuff.cpp:49-56: ORB::create()->detectAndCompute feeds
// DBoW2).  SURVEY.md section 8(f)-2.  Every stage is bit-identical to cv2 4.13.0 (oracle/orb.py is the stage-by-stage
// restatement): the INTER_LINEAR_EXACT pyramid, FAST-9/16 with suppression, the Harris ranking response, IC_Angle,
// the smoothing ORB really applies, the rBRIEF descriptors (test-pair table recovered from cv2 itself, orb_pattern.h).
// vo_orb_detect_and_compute runs them per level with OpenCV's selection rules (quota, border filter, retainBest)
// on the host between the kernels.
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
  uint8_t* h_pin = nullptr;        // pinned host scratch of vo_orb_detect_and_compute: image | xy | scalar | descriptors
  size_t h_pin_bytes = 0;
};

__constant__ signed char c_orb_pattern[256][4];
__constant__ float c_orb_gauss[4];

void orb_free(vo_ctx* c) {
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  if (!o) return;
  void* dev[] = {o->img, o->rowf, o->sm, o->xy, o->ang, o->desc, o->score, o->flag, o->sel, o->d_n, o->cub_tmp, o->pyr, o->coef};
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

// what one pyramid level contributes to the result
struct OrbLevelOut {
  int level = 0, n = 0, pin_ofs = 0;     // pin_ofs: where its angles / descriptors sit in the pinned result staging
  std::vector<float> xy, resp;
};

// KeyPointsFilter::retainBest: every keypoint whose response is >= the n-th largest stays (ties included)
std::vector<int> orb_retain_best(const std::vector<float>& resp, int n) {
  std::vector<int> keep;
  if (n <= 0) return keep;
  if ((int)resp.size() <= n) {
    keep.resize(resp.size());
    std::iota(keep.begin(), keep.end(), 0);
    return keep;
  }
  std::vector<float> tmp(resp);
  std::nth_element(tmp.begin(), tmp.begin() + (n - 1), tmp.end(), std::greater<float>());
  const float thr = tmp[n - 1];
  for (int i = 0; i < (int)resp.size(); i++)
    if (resp[i] >= thr) keep.push_back(i);
  return keep;
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
  constexpr int NL = 8, EDGE = 31, FAST_T = 20;
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
  bool active[NL];
  int cand_ofs[NL + 1];
  cand_ofs[0] = 0;
  for (int l = 0; l < NL; l++) {
    active[l] = lw[l] > 2 * EDGE && lh[l] > 2 * EDGE && quota[l] > 0;      // else the border filter leaves nothing
    cand_ofs[l + 1] = cand_ofs[l] + (active[l] ? lw[l] * lh[l] / 4 + 16 : 0);
  }
  const int n_cand = std::max(cand_ofs[NL], 1);
  VO_TRY(orb_ensure(c, width, height, n_cand));
  Orb* o = reinterpret_cast<Orb*>(c->orb);
  VO_TRY(orb_ensure_fast(c, o));
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
  std::vector<int> h_coef;
  size_t coef_ofs[NL] = {0};
  for (int l = 1; l < NL; l++) {
    coef_ofs[l] = h_coef.size();
    h_coef.resize(h_coef.size() + 2 * (size_t)(lw[l] + lh[l]));
    int* t = h_coef.data() + coef_ofs[l];
    orb_exact_coeffs(lw[l - 1], lw[l], t, t + lw[l]);
    orb_exact_coeffs(lh[l - 1], lh[l], t + 2 * lw[l], t + 2 * lw[l] + lh[l]);
  }
  if ((int)h_coef.size() > o->coef_cap) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(o->coef);
    o->coef = nullptr;
    VO_CUDA(cudaMalloc(&o->coef, h_coef.size() * sizeof(int)));
    o->coef_cap = (int)h_coef.size();
  }
  // pinned scratch: pageable copies are staged by the driver at a few GB/s and serialise with the stream
  const int DESC_CAP = std::max(8192, 2 * nfeatures + 4096);     // keypoints the result staging holds
  // every section starts 16-byte aligned (float views, aligned async copies)
  const size_t pin_img = (((size_t)width * height) + 15) & ~(size_t)15, pin_xy = 2 * (size_t)o->cap * sizeof(float),
               pin_sc = (size_t)o->cap * sizeof(float), pin_desc = (size_t)DESC_CAP * (32 + sizeof(float)) + 64;
  if (pin_img + pin_xy + pin_sc + pin_desc > o->h_pin_bytes) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    cudaFreeHost(o->h_pin);
    o->h_pin = nullptr;
    o->h_pin_bytes = 0;
    VO_CUDA(cudaMallocHost(&o->h_pin, pin_img + pin_xy + pin_sc + pin_desc));
    o->h_pin_bytes = pin_img + pin_xy + pin_sc + pin_desc;
  }
  float* p_xy = reinterpret_cast<float*>(o->h_pin + pin_img);
  float* p_sc = reinterpret_cast<float*>(o->h_pin + pin_img + pin_xy);
  uint8_t* p_desc = o->h_pin + pin_img + pin_xy + pin_sc;                    // descriptors of all levels, appended
  float* p_ang = reinterpret_cast<float*>(p_desc + (size_t)DESC_CAP * 32);   // their angles
  for (int y = 0; y < height; y++) memcpy(o->h_pin + (size_t)y * width, img + (size_t)y * stride, width);
  VO_CUDA(cudaMemcpyAsync(o->img, o->h_pin, (size_t)width * height, cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(o->coef, h_coef.data(), h_coef.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  const uint8_t* level[NL];
  level[0] = o->img;
  {
    uint8_t* next = o->pyr;
    for (int l = 1; l < NL; l++) {
      const int* t = o->coef + coef_ofs[l];
      LaunchScope ls(c, VO_K_MISC);
      orb_resize_exact_kernel<<<dim3(div_up(lw[l], 128), lh[l]), 128, 0, c->stream>>>(level[l - 1], lw[l - 1], lh[l - 1], next, lw[l],
                                                                                    lh[l], t, t + lw[l], t + 2 * lw[l],
                                                                                    t + 2 * lw[l] + lh[l]);
      level[l] = next;
      next += (size_t)lw[l] * lh[l];
    }
  }

  // ---- phase 1: FAST (threshold 20, suppression) on every level, corners and scores into the level's slice
  for (int l = 0; l < NL; l++) {
    if (!active[l]) continue;
    const int w = lw[l], h = lh[l];
    {
      LaunchScope ls(c, VO_K_MISC);
      fast_score_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(level[l], w, h, FAST_T, o->score);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      fast_nms_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(o->score, w, h, 1, o->flag);
    }
    size_t tb = o->cub_bytes;
    VO_CUDA(cub::DeviceSelect::Flagged(o->cub_tmp, tb, cub::CountingInputIterator<int>(0), o->flag, o->sel, o->d_n + l, w * h,
                                       c->stream));
    c->launch_count++;
    const int ccap = cand_ofs[l + 1] - cand_ofs[l];
    {
      LaunchScope ls(c, VO_K_MISC);
      fast_gather_kernel<<<div_up(ccap, 256), 256, 0, c->stream>>>(o->sel, o->d_n + l, ccap, o->score, w, o->xy + 2 * cand_ofs[l],
                                                                   o->ang + cand_ofs[l]);
    }
  }
  VO_CUDA(cudaGetLastError());
  int* p_cnt = reinterpret_cast<int*>(p_ang + DESC_CAP);
  VO_CUDA(cudaMemcpyAsync(p_cnt, o->d_n, NL * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  int nc[NL];
  for (int l = 0; l < NL; l++) {
    nc[l] = active[l] ? std::min(p_cnt[l], cand_ofs[l + 1] - cand_ofs[l]) : 0;
    if (nc[l] == 0) continue;
    VO_CUDA(cudaMemcpyAsync(p_xy + 2 * cand_ofs[l], o->xy + 2 * cand_ofs[l], 2 * (size_t)nc[l] * sizeof(float),
                            cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(p_sc + cand_ofs[l], o->ang + cand_ofs[l], (size_t)nc[l] * sizeof(float), cudaMemcpyDeviceToHost,
                            c->stream));
  }
  VO_CUDA(cudaStreamSynchronize(c->stream));

  // ---- phase 2: KeyPointsFilter::runByImageBorder(edgeThreshold), retainBest(2 * featuresNum) by FAST score, then
  // HarrisResponses of the survivors on every level
  std::vector<float> sxy[NL];
  int n1[NL];
  for (int l = 0; l < NL; l++) {
    n1[l] = 0;
    if (nc[l] == 0) continue;
    const int w = lw[l], h = lh[l];
    const float* h_xy = p_xy + 2 * cand_ofs[l];
    const float* h_sc = p_sc + cand_ofs[l];
    std::vector<float> kxy, ksc;
    for (int i = 0; i < nc[l]; i++) {
      const float x = h_xy[2 * i], y = h_xy[2 * i + 1];
      if (x >= EDGE && x < w - EDGE && y >= EDGE && y < h - EDGE) {
        kxy.push_back(x);
        kxy.push_back(y);
        ksc.push_back(h_sc[i]);
      }
    }
    const std::vector<int> keep = orb_retain_best(ksc, 2 * quota[l]);
    for (int i : keep) {
      sxy[l].push_back(kxy[2 * i]);
      sxy[l].push_back(kxy[2 * i + 1]);
    }
    n1[l] = (int)keep.size();
  }
  for (int l = 0; l < NL; l++) {
    if (n1[l] == 0) continue;
    float* px = p_xy + 2 * cand_ofs[l];            // the phase-1 results in this slice have been consumed
    memcpy(px, sxy[l].data(), sxy[l].size() * sizeof(float));
    VO_CUDA(cudaMemcpyAsync(o->xy + 2 * cand_ofs[l], px, sxy[l].size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    {
      LaunchScope ls(c, VO_K_MISC);
      orb_harris_kernel<<<div_up(n1[l] * 32, 256), 256, 0, c->stream>>>(level[l], lw[l], lh[l], o->xy + 2 * cand_ofs[l], n1[l], 0.04f,
                                                                       o->ang + cand_ofs[l]);
    }
    VO_CUDA(cudaMemcpyAsync(p_sc + cand_ofs[l], o->ang + cand_ofs[l], (size_t)n1[l] * sizeof(float), cudaMemcpyDeviceToHost,
                            c->stream));
  }
  VO_CUDA(cudaGetLastError());
  VO_CUDA(cudaStreamSynchronize(c->stream));

  // ---- phase 3: retainBest(featuresNum) by Harris response, then ICAngles on the unsmoothed level, GaussianBlur of the
  // level and computeOrbDescriptors; the order inside a level is raster (y, x), OpenCV's is what std::nth_element leaves
  std::vector<OrbLevelOut> pending;
  int pin_used = 0;
  for (int l = 0; l < NL; l++) {
    if (n1[l] == 0) continue;
    const int w = lw[l], h = lh[l];
    const float* resp = p_sc + cand_ofs[l];
    const std::vector<float> h_resp(resp, resp + n1[l]);
    std::vector<int> keep = orb_retain_best(h_resp, quota[l]);
    const std::vector<float>& sx = sxy[l];
    std::sort(keep.begin(), keep.end(), [&](int a, int b) {
      return sx[2 * a + 1] != sx[2 * b + 1] ? sx[2 * a + 1] < sx[2 * b + 1] : sx[2 * a] < sx[2 * b];
    });
    const int n2 = (int)keep.size();
    if (n2 == 0) continue;
    if (pin_used + n2 > DESC_CAP) {
      set_error("vo_orb_detect_and_compute: more than %d keypoints", DESC_CAP);
      cudaStreamSynchronize(c->stream);   // kernels and copies into the pinned staging of earlier levels are still in flight
      return VO_ERR_CAPACITY;
    }
    OrbLevelOut rec;
    rec.level = l;
    rec.n = n2;
    rec.pin_ofs = pin_used;
    rec.xy.resize(2 * (size_t)n2);
    rec.resp.resize(n2);
    for (int i = 0; i < n2; i++) {
      rec.xy[2 * i] = sx[2 * keep[i]];
      rec.xy[2 * i + 1] = sx[2 * keep[i] + 1];
      rec.resp[i] = h_resp[keep[i]];
    }
    float* px = p_xy + 2 * cand_ofs[l];
    memcpy(px, rec.xy.data(), rec.xy.size() * sizeof(float));
    float* d_xy = o->xy + 2 * cand_ofs[l];
    float* d_ang = o->ang + cand_ofs[l];
    uint8_t* d_desc = o->desc + (size_t)cand_ofs[l] * 32;
    VO_CUDA(cudaMemcpyAsync(d_xy, px, rec.xy.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    {
      LaunchScope ls(c, VO_K_MISC);
      orb_angle_kernel<<<div_up(n2 * 32, 256), 256, 0, c->stream>>>(level[l], w, h, d_xy, n2, d_ang);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      orb_smooth_rows_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(level[l], w, h, o->rowf);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      orb_smooth_cols_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(o->rowf, w, h, o->sm);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      orb_describe_kernel<<<div_up(n2 * 32, 256), 256, 0, c->stream>>>(o->sm, w, h, d_xy, d_ang, n2, d_desc);
    }
    VO_CUDA(cudaMemcpyAsync(p_ang + pin_used, d_ang, (size_t)n2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(p_desc + (size_t)pin_used * 32, d_desc, (size_t)n2 * 32, cudaMemcpyDeviceToHost, c->stream));
    pin_used += n2;
    pending.push_back(std::move(rec));
  }
  VO_CUDA(cudaGetLastError());
  VO_CUDA(cudaStreamSynchronize(c->stream));
  int total = 0;
  for (const auto& r : pending) {
    const float* ang = p_ang + r.pin_ofs;
    const uint8_t* dsc = p_desc + (size_t)r.pin_ofs * 32;
    for (int i = 0; i < r.n; i++) {
      if (total < cap) {
        xy[2 * total] = r.xy[2 * i] * lscale[r.level];          // allKeypoints[i].pt *= scale
        xy[2 * total + 1] = r.xy[2 * i + 1] * lscale[r.level];
        if (octave) octave[total] = r.level;
        if (response) response[total] = r.resp[i];
        if (angle_deg) angle_deg[total] = ang[i];
        memcpy(desc + (size_t)total * 32, dsc + (size_t)i * 32, 32);
      }
      total++;
    }
  }
  *n_out = total;
  return total > cap ? VO_ERR_CAPACITY : VO_OK;
}
