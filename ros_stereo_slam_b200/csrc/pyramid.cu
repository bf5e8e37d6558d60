// pyramid.cu -- K1: image pyramid (pyrDown x3) + Scharr derivative, the part of
// cv::calcOpticalFlowPyrLK (reference call sites src/tracking.cpp:18,52) that
// cv::buildOpticalFlowPyramid performs.  Integer arithmetic, bit-exact with OpenCV:
//   pyrDown : separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8
//   Scharr  : dx = S(x+1)-S(x-1), S = 3*(I(y-1)+I(y+1)) + 10*I(y);
//             dy = 3*(D(x-1)+D(x+1)) + 10*D(x), D = I(y+1)-I(y-1)   (gain 32, int16)
// Levels live PADDED in HBM (common.cuh) so neither this file's taps nor the LK window
// need bounds logic: image borders = REFLECT_101, derivative borders = 0.
#include "common.cuh"

namespace vo {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Fill the REFLECT_101 border of a padded level whose interior is already in place.
// One thread per 4 padded pixels of the border rows / border columns.
__global__ void border_kernel(uint8_t* __restrict__ img, int w, int h, int pitch) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;  // padded x
  const int py = blockIdx.y;                             // padded y
  const int pw = w + PAD_L + PAD_R;
  if (px >= pw) return;
  const int x = px - PAD_L, y = py - PAD_Y;
  if (x >= 0 && x < w && y >= 0 && y < h) return;  // interior
  // far-out columns (beyond what reflect101 can address) are clamped; they are never read
  int xr = x < -(w - 1) ? 0 : x > 2 * (w - 1) ? w - 1 : reflect101(x, w);
  int yr = reflect101(y, h);
  img[(size_t)py * pitch + px] = img[(size_t)(yr + PAD_Y) * pitch + xr + PAD_L];
}

// pyrDown: dst padded level (including its border) from src padded level.
// Each thread produces one dst pixel; 5x5 taps read through the read-only path.
__global__ void pyrdown_kernel(const uint8_t* __restrict__ src, int spitch, uint8_t* __restrict__ dst, int dw, int dh,
                               int dpitch) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y * blockDim.y + threadIdx.y;
  const int pw = dw + PAD_L + PAD_R, ph = dh + 2 * PAD_Y;
  if (px >= pw || py >= ph) return;
  int x = px - PAD_L, y = py - PAD_Y;
  x = x < -(dw - 1) ? 0 : x > 2 * (dw - 1) ? dw - 1 : reflect101(x, dw);
  y = reflect101(y, dh);
  const uint8_t* s = src + (size_t)(2 * y - 2 + PAD_Y) * spitch + (2 * x - 2 + PAD_L);
  int acc = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int kj = (j == 0 || j == 4) ? 1 : (j == 2 ? 6 : 4);
    const uint8_t* r = s + (size_t)j * spitch;
    int row = (int)__ldg(r) + 4 * (int)__ldg(r + 1) + 6 * (int)__ldg(r + 2) + 4 * (int)__ldg(r + 3) + (int)__ldg(r + 4);
    acc += kj * row;
  }
  dst[(size_t)py * dpitch + px] = (uint8_t)((acc + 128) >> 8);
}

// Scharr derivative of a padded level: interior from the (reflect-padded) image, border 0.
__global__ void scharr_kernel(const uint8_t* __restrict__ img, short2* __restrict__ deriv, int w, int h, int pitch) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y * blockDim.y + threadIdx.y;
  const int pw = w + PAD_L + PAD_R, ph = h + 2 * PAD_Y;
  if (px >= pw || py >= ph) return;
  const int x = px - PAD_L, y = py - PAD_Y;
  short2 out = make_short2(0, 0);
  if (x >= 0 && x < w && y >= 0 && y < h) {
    const uint8_t* p = img + (size_t)py * pitch + px;
    int a00 = __ldg(p - pitch - 1), a01 = __ldg(p - pitch), a02 = __ldg(p - pitch + 1);
    int a10 = __ldg(p - 1), a12 = __ldg(p + 1);
    int a20 = __ldg(p + pitch - 1), a21 = __ldg(p + pitch), a22 = __ldg(p + pitch + 1);
    int sl = 3 * (a00 + a20) + 10 * a10;   // vertical smoothing at x-1
    int sr = 3 * (a02 + a22) + 10 * a12;   // at x+1
    int dl = a20 - a00, dc = a21 - a01, dr = a22 - a02;  // vertical derivative at x-1, x, x+1
    out.x = (short)(sr - sl);
    out.y = (short)(3 * (dl + dr) + 10 * dc);
  }
  deriv[(size_t)py * pitch + px] = out;
}

// ------------------------------------------------------------------------------------
int pyr_alloc(vo_ctx* c, Pyramid& p) {
  int w = c->p.width, h = c->p.height;
  p.nlevels = 0;
  for (int l = 0; l <= c->p.lk_max_level && l < MAX_LEVELS; l++) {
    if (l > 0) {
      int nw = (w + 1) / 2, nh = (h + 1) / 2;
      if (nw <= c->p.lk_win || nh <= c->p.lk_win) break;  // buildOpticalFlowPyramid stops here
      w = nw;
      h = nh;
    }
    PyrLevel& L = p.lv[l];
    L.w = w;
    L.h = h;
    L.pitch = ((w + PAD_L + PAD_R + 127) / 128) * 128;
    size_t rows = (size_t)h + 2 * PAD_Y;
    VO_CUDA(cudaMalloc(&L.img, rows * L.pitch));
    VO_CUDA(cudaMalloc(&L.deriv, rows * L.pitch * sizeof(short2)));
    VO_CUDA(cudaMemsetAsync(L.img, 0, rows * L.pitch, c->stream));
    VO_CUDA(cudaMemsetAsync(L.deriv, 0, rows * L.pitch * sizeof(short2), c->stream));
    p.nlevels = l + 1;
  }
  p.has_deriv = false;
  p.stamp = 0;
  return VO_OK;
}

void pyr_free(Pyramid& p) {
  for (int l = 0; l < p.nlevels; l++) {
    cudaFree(p.lv[l].img);
    cudaFree(p.lv[l].deriv);
    p.lv[l].img = nullptr;
    p.lv[l].deriv = nullptr;
  }
  p.nlevels = 0;
}

PyrView pyr_view(const Pyramid& p) {
  PyrView v;
  v.nlevels = p.nlevels;
  for (int l = 0; l < MAX_LEVELS; l++) {
    if (l < p.nlevels) {
      v.lv[l].img = p.lv[l].img;
      v.lv[l].deriv = p.lv[l].deriv;
      v.lv[l].w = p.lv[l].w;
      v.lv[l].h = p.lv[l].h;
      v.lv[l].pitch = p.lv[l].pitch;
    } else {
      v.lv[l] = PyrLevelView{nullptr, nullptr, 0, 0, 0};
    }
  }
  return v;
}

static int scharr_all(vo_ctx* c, Pyramid& p) {
  for (int l = 0; l < p.nlevels; l++) {
    PyrLevel& L = p.lv[l];
    dim3 b(32, 8), g(div_up(L.w + PAD_L + PAD_R, 32), div_up(L.h + 2 * PAD_Y, 8));
    LaunchScope ls(c, VO_K_PYRAMID);
    scharr_kernel<<<g, b, 0, c->stream>>>(L.img, L.deriv, L.w, L.h, L.pitch);
  }
  VO_CUDA(cudaGetLastError());
  p.has_deriv = true;
  return VO_OK;
}

// Level 0 interior must already be in place (memcpy2D straight into the padded buffer).
int pyr_build(vo_ctx* c, int slot, const uint8_t* d_tight, bool with_deriv) {
  Pyramid& p = c->pyr[slot];
  PyrLevel& L0 = p.lv[0];
  if (d_tight) {
    VO_CUDA(cudaMemcpy2DAsync(L0.img + (size_t)PAD_Y * L0.pitch + PAD_L, L0.pitch, d_tight, L0.w, L0.w, L0.h,
                              cudaMemcpyDeviceToDevice, c->stream));
  }
  {
    dim3 b(128), g(div_up(L0.w + PAD_L + PAD_R, 128), L0.h + 2 * PAD_Y);
    LaunchScope ls(c, VO_K_PYRAMID);
    border_kernel<<<g, b, 0, c->stream>>>(L0.img, L0.w, L0.h, L0.pitch);
  }
  for (int l = 1; l < p.nlevels; l++) {
    PyrLevel& S = p.lv[l - 1];
    PyrLevel& D = p.lv[l];
    dim3 b(32, 8), g(div_up(D.w + PAD_L + PAD_R, 32), div_up(D.h + 2 * PAD_Y, 8));
    LaunchScope ls(c, VO_K_PYRAMID);
    pyrdown_kernel<<<g, b, 0, c->stream>>>(S.img, S.pitch, D.img, D.w, D.h, D.pitch);
  }
  VO_CUDA(cudaGetLastError());
  p.has_deriv = false;
  p.stamp = ++c->stamp_counter;
  if (with_deriv) return scharr_all(c, p);
  return VO_OK;
}

int pyr_ensure_deriv(vo_ctx* c, int slot) {
  Pyramid& p = c->pyr[slot];
  if (p.has_deriv) return VO_OK;
  return scharr_all(c, p);
}

}  // namespace vo
