// pyramid.cu -- K1: image pyramid (pyrDown x3) + Scharr derivative, the part of
// cv::calcOpticalFlowPyrLK (reference call sites src/tracking.cpp:18,52) that
// cv::buildOpticalFlowPyramid performs.  Integer arithmetic, bit-exact with OpenCV:
//   pyrDown : separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8
//   Scharr  : dx = S(x+1)-S(x-1), S = 3*(I(y-1)+I(y+1)) + 10*I(y);
//             dy = 3*(D(x-1)+D(x+1)) + 10*D(x), D = I(y+1)-I(y-1)   (gain 32, int16)
// Levels live PADDED in HBM (common.cuh) so the LK window needs no bounds logic: image
// borders = REFLECT_101, derivative borders = 0.
//
// ONE launch per image (pyr_fused_kernel): a CTA owns a 64x32 block of level 0 and the
// blocks it maps to on levels 1..3 (32x16, 16x8, 8x4).  Its level-0 tile including the halo
// every deeper level needs (101 x 69 pixels, box 128 x 69) is fetched by a single TMA box load
// (cp.async.bulk.tensor.2d -> shared memory, mbarrier completion; the tensor map covers the
// level-0 interior, so out-of-image parts of the box are zero-filled and never read);
// levels 1..3 are then produced tile by tile in shared memory (49x33, 23x15, 10x6), each from
// the tile above it with REFLECT_101 applied to the tap coordinates -- the same values OpenCV
// computes level by level -- and the owned pixels, their mirror images in the padded border and
// the Scharr derivatives of all four levels are written with 4-pixel (uchar4 / 2 x short2x2)
// vector stores.  8 launches + 4 lazy Scharr launches of the first version -> 1.
#include <cuda.h>

#include "common.cuh"

namespace vo {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Scharr derivative of a padded level: interior from the (reflect-padded) image, border 0.
// Only used when a pyramid that was built without derivatives later serves as the previous image.
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

// ------------------------------------------------------------------------------------ fused K1
constexpr int K1_T0W = 64, K1_T0H = 32;      // level-0 block owned by one CTA
constexpr int K1_BOXW = 128, K1_BOXH = 69;   // TMA box: 101 x 69 needed; the x origin is rounded down to 16 B
                                             // (a TMA box must start 16-byte aligned in the inner dimension: +15)
constexpr int K1_THREADS = 256;
// tile capacities of levels 1..3 (owned block + halo of everything deeper, see need_range)
constexpr int K1_TW1 = 49, K1_TH1 = 33, K1_TW2 = 23, K1_TH2 = 15, K1_TW3 = 10, K1_TH3 = 6;

struct K1Level {
  uint8_t* img;
  short2* deriv;
  int w, h, pitch;
};
struct K1Args {
  K1Level lv[MAX_LEVELS];       // plane 0; plane z at img + z*plane[l], deriv + z*plane[l]
  size_t plane[MAX_LEVELS];
  int nlevels;
  int with_deriv;
};
struct K1Maps {
  CUtensorMap m[MAX_CN];        // level-0 interior of each plane
};

struct Range {
  int lo, hi;   // inclusive, inside the image
};

// In-image coordinates that REFLECT_101 maps [lo, hi] onto (lo may be < 0, hi may be >= n).
__device__ __forceinline__ Range reflect_closure(int lo, int hi, int n) {
  Range r;
  r.lo = max(lo, 0);
  r.hi = min(hi, n - 1);
  if (lo < 0) r.hi = max(r.hi, min(-lo, n - 1));
  if (hi > n - 1) r.lo = min(r.lo, max(2 * (n - 1) - hi, 0));
  return r;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// dst tile (coordinates dx x dy of the next level) <- pyrDown of the src tile
__device__ __forceinline__ void k1_pyrdown_tile(const uint8_t* __restrict__ src, int sx0, int sy0, int sstride, int sw,
                                                int sh, uint8_t* __restrict__ dst, Range dx, Range dy, int dstride) {
  const int W = dx.hi - dx.lo + 1, H = dy.hi - dy.lo + 1;
  for (int i = threadIdx.x; i < W * H; i += K1_THREADS) {
    const int y = i / W, x = i - y * W;
    const int X = 2 * (dx.lo + x), Y = 2 * (dy.lo + y);
    int xo[5];
#pragma unroll
    for (int k = 0; k < 5; k++) xo[k] = reflect101(X + k - 2, sw) - sx0;
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      const int kj = (j == 0 || j == 4) ? 1 : (j == 2 ? 6 : 4);
      const uint8_t* r = src + (reflect101(Y + j - 2, sh) - sy0) * sstride;
      acc += kj * ((int)r[xo[0]] + 4 * (int)r[xo[1]] + 6 * (int)r[xo[2]] + 4 * (int)r[xo[3]] + (int)r[xo[4]]);
    }
    dst[y * dstride + x] = (uint8_t)((acc + 128) >> 8);
  }
}

// Global writes of one level from its shared-memory tile: owned pixels (levels >= 1; level 0's
// interior is already in place), their REFLECT_101 mirror images in the padded border, and the
// Scharr derivative of the owned pixels.  4 pixels per thread, vector stores.
__device__ __forceinline__ void k1_emit_level(const uint8_t* __restrict__ tile, int tx0, int ty0, int tstride,
                                              const K1Level L, Range ox, Range oy, bool write_interior, bool with_deriv) {
  const int w = L.w, h = L.h;
  const int OW = ox.hi - ox.lo + 1, OH = oy.hi - oy.lo + 1;
  const int G = (OW + 3) >> 2;                  // 4-pixel groups per row (ox.lo is a multiple of 8)
  for (int i = threadIdx.x; i < G * OH; i += K1_THREADS) {
    const int gy = i / G, gx = i - gy * G;
    const int X = ox.lo + 4 * gx, Y = oy.lo + gy;
    const int nv = min(4, ox.hi + 1 - X);
    const uint8_t* rc = tile + (Y - ty0) * tstride - tx0;      // row Y of the tile, indexed by image x
    if (write_interior) {
      uint8_t* d = L.img + (size_t)(Y + PAD_Y) * L.pitch + PAD_L + X;
      if (nv == 4) {
        *reinterpret_cast<uchar4*>(d) = make_uchar4(rc[X], rc[X + 1], rc[X + 2], rc[X + 3]);
      } else {
        for (int k = 0; k < nv; k++) d[k] = rc[X + k];
      }
    }
    if (with_deriv) {
      const uint8_t* ru = tile + (reflect101(Y - 1, h) - ty0) * tstride - tx0;
      const uint8_t* rd = tile + (reflect101(Y + 1, h) - ty0) * tstride - tx0;
      int cu[6], cc[6], cd[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int xx = reflect101(min(X + k - 1, w), w);      // min(): columns past the last group are not used
        cu[k] = ru[xx];
        cc[k] = rc[xx];
        cd[k] = rd[xx];
      }
      short2 o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int sl = 3 * (cu[k] + cd[k]) + 10 * cc[k];            // vertical smoothing at x-1
        const int sr = 3 * (cu[k + 2] + cd[k + 2]) + 10 * cc[k + 2];  // at x+1
        const int dl = cd[k] - cu[k], dc = cd[k + 1] - cu[k + 1], dr = cd[k + 2] - cu[k + 2];
        o[k] = make_short2((short)(sr - sl), (short)(3 * (dl + dr) + 10 * dc));
      }
      short2* d = L.deriv + (size_t)(Y + PAD_Y) * L.pitch + PAD_L + X;
      if (nv == 4) {
        int4 v;
        v.x = *reinterpret_cast<int*>(&o[0]);
        v.y = *reinterpret_cast<int*>(&o[1]);
        v.z = *reinterpret_cast<int*>(&o[2]);
        v.w = *reinterpret_cast<int*>(&o[3]);
        *reinterpret_cast<int4*>(d) = v;
      } else {
        for (int k = 0; k < nv; k++) d[k] = o[k];
      }
    }
  }
  // mirror images: pixel x also lands on -x (1 <= x <= PAD) and on 2(w-1)-x (w-1-PAD <= x <= w-2)
  const int mxl = min(PAD_L, w - 1), mxr = min(PAD_R, w - 1), my = min(PAD_Y, h - 1);
  const bool edge_x = ox.lo <= mxl || ox.hi >= w - 1 - mxr;
  const bool edge_y = oy.lo <= my || oy.hi >= h - 1 - my;
  if (!edge_x && !edge_y) return;
  for (int i = threadIdx.x; i < OW * OH; i += K1_THREADS) {
    const int yy = i / OW, xx = i - yy * OW;
    const int X = ox.lo + xx, Y = oy.lo + yy;
    const uint8_t v = tile[(Y - ty0) * tstride + (X - tx0)];
    int tx[3], ty[3], nx = 1, ny = 1;
    tx[0] = X;
    ty[0] = Y;
    if (X >= 1 && X <= mxl) tx[nx++] = -X;
    if (X <= w - 2 && X >= w - 1 - mxr) tx[nx++] = 2 * (w - 1) - X;
    if (Y >= 1 && Y <= my) ty[ny++] = -Y;
    if (Y <= h - 2 && Y >= h - 1 - my) ty[ny++] = 2 * (h - 1) - Y;
    for (int a = 0; a < ny; a++)
      for (int b = 0; b < nx; b++) {
        if (a == 0 && b == 0) continue;
        L.img[(size_t)(ty[a] + PAD_Y) * L.pitch + PAD_L + tx[b]] = v;
      }
  }
}

__global__ void __launch_bounds__(K1_THREADS)
pyr_fused_kernel(const __grid_constant__ K1Maps maps, const K1Args a) {
  __shared__ alignas(128) uint8_t t0[K1_BOXW * K1_BOXH];
  __shared__ uint8_t t1[K1_TW1 * K1_TH1], t2[K1_TW2 * K1_TH2], t3[K1_TW3 * K1_TH3];
  __shared__ alignas(8) unsigned long long bar;

  // ---- coordinate ranges: owned block and needed (halo) block per level, deepest level first
  Range ox[MAX_LEVELS], oy[MAX_LEVELS], nx[MAX_LEVELS], ny[MAX_LEVELS];
  const int nl = a.nlevels;
#pragma unroll
  for (int l = MAX_LEVELS - 1; l >= 0; l--) {
    if (l >= nl) continue;
    const int w = a.lv[l].w, h = a.lv[l].h;
    ox[l].lo = (blockIdx.x * K1_T0W) >> l;
    ox[l].hi = min(ox[l].lo + (K1_T0W >> l) - 1, w - 1);
    oy[l].lo = (blockIdx.y * K1_T0H) >> l;
    oy[l].hi = min(oy[l].lo + (K1_T0H >> l) - 1, h - 1);
    // Scharr (and nothing else) reads one pixel around the owned block ...
    int xlo = ox[l].lo - 1, xhi = ox[l].hi + 1, ylo = oy[l].lo - 1, yhi = oy[l].hi + 1;
    if (l + 1 < nl) {   // ... and pyrDown of the next level's tile reads 2x-2 .. 2x+2
      xlo = min(xlo, 2 * nx[l + 1].lo - 2);
      xhi = max(xhi, 2 * nx[l + 1].hi + 2);
      ylo = min(ylo, 2 * ny[l + 1].lo - 2);
      yhi = max(yhi, 2 * ny[l + 1].hi + 2);
    }
    nx[l] = reflect_closure(xlo, xhi, w);
    ny[l] = reflect_closure(ylo, yhi, h);
  }

  // ---- level-0 tile: one TMA box load
  const int t0x = nx[0].lo & ~15;
  const uint32_t bar_a = smem_addr(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(K1_BOXW * K1_BOXH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_addr(t0)),
        "l"(reinterpret_cast<uint64_t>(&maps.m[blockIdx.z])), "r"(t0x), "r"(ny[0].lo), "r"(bar_a)
        : "memory");
  }
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "K1_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra K1_DONE;\n"
      "bra K1_WAIT;\n"
      "K1_DONE:\n"
      "}\n" ::"r"(bar_a),
      "r"(0)
      : "memory");

  // ---- levels 1..3 in shared memory
  uint8_t* tiles[MAX_LEVELS] = {t0, t1, t2, t3};
  const int strides[MAX_LEVELS] = {K1_BOXW, K1_TW1, K1_TW2, K1_TW3};
#pragma unroll
  for (int l = 1; l < MAX_LEVELS; l++) {
    if (l >= nl) break;
    k1_pyrdown_tile(tiles[l - 1], l == 1 ? t0x : nx[l - 1].lo, ny[l - 1].lo, strides[l - 1], a.lv[l - 1].w, a.lv[l - 1].h, tiles[l],
                    nx[l], ny[l], strides[l]);
    __syncthreads();
  }
  // ---- global writes
#pragma unroll
  for (int l = 0; l < MAX_LEVELS; l++) {
    if (l >= nl) break;
    K1Level L = a.lv[l];
    L.img += blockIdx.z * a.plane[l];
    L.deriv += blockIdx.z * a.plane[l];
    k1_emit_level(tiles[l], l == 0 ? t0x : nx[l].lo, ny[l].lo, strides[l], L, ox[l], oy[l], l > 0, a.with_deriv != 0);
  }
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

// Interleaved BGR (tight rows of 3*w bytes) -> the level-0 interiors of the three planes.
// 4 pixels per thread: twelve byte loads (rows of 3*w bytes are not word aligned), three uchar4 stores.
// Also clears *mono (preset to 1) when any pixel has B != G or G != R: a gray image that imread returned as BGR --
// the reference's KITTI case -- has three identical planes, which the LK launch exploits (lk.cu, lk_kernel<.., true>).
__global__ void split_planes_kernel(const uint8_t* __restrict__ bgr, int w, int h, uint8_t* __restrict__ img, int pitch,
                                    size_t plane, int* __restrict__ mono) {
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* s = bgr + (size_t)y * (3 * w) + 3 * x;
  uint8_t* d = img + (size_t)(y + PAD_Y) * pitch + PAD_L + x;
  bool same = true;
  if (x + 4 <= w) {
    uint8_t v[12];
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = __ldg(s + k);
#pragma unroll
    for (int c = 0; c < 3; c++)
      *reinterpret_cast<uchar4*>(d + c * plane) = make_uchar4(v[c], v[3 + c], v[6 + c], v[9 + c]);
#pragma unroll
    for (int k = 0; k < 4; k++) same = same && v[3 * k] == v[3 * k + 1] && v[3 * k] == v[3 * k + 2];
  } else {
    for (int k = 0; x + k < w; k++) {
      for (int c = 0; c < 3; c++) d[c * plane + k] = s[3 * k + c];
      same = same && s[3 * k] == s[3 * k + 1] && s[3 * k] == s[3 * k + 2];
    }
  }
  if (!same) *mono = 0;     // benign race: every writer stores 0
}

// Device image with an arbitrary row stride -> level-0 interior of a padded plane.  An SM copy instead of
// cudaMemcpy2DAsync(DeviceToDevice): the copy engines stay free for the host<->device transfers that run
// beside the frame (vo_seq_prefetch), and 1241-byte rows are 376 separate DMA rows for them.
__global__ void unpack_rows_kernel(const uint8_t* __restrict__ src, int w, int h, int src_pitch, uint8_t* __restrict__ img,
                                   int pitch) {
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* s = src + (size_t)y * src_pitch + x;
  uint8_t* d = img + (size_t)(y + PAD_Y) * pitch + PAD_L + x;
  if (x + 4 <= w) {
    *reinterpret_cast<uchar4*>(d) = make_uchar4(__ldg(s), __ldg(s + 1), __ldg(s + 2), __ldg(s + 3));
  } else {
    for (int k = 0; x + k < w; k++) d[k] = s[k];
  }
}

int pyr_unpack_rows(vo_ctx* c, int slot, const uint8_t* d_src, int src_pitch) {
  PyrLevel& L0 = c->pyr[slot].lv[0];
  dim3 b(128), g(div_up(div_up(L0.w, 4), 128), L0.h);
  {
    LaunchScope ls(c, VO_K_PYRAMID);
    unpack_rows_kernel<<<g, b, 0, c->stream>>>(d_src, L0.w, L0.h, src_pitch, L0.img, L0.pitch);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// cv::cvtColor(BGR2GRAY) for 8-bit images (reference src/StereoCV.cpp:35-36): OpenCV's 15-bit fixed point
// gray = (B*3735 + G*19235 + R*9798 + 2^14) >> 15.  4 pixels per thread, one uchar4 store.
__global__ void bgr2gray_kernel(const uint8_t* __restrict__ bgr, int w, int h, int src_pitch, uint8_t* __restrict__ gray,
                                int dst_pitch) {
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* s = bgr + (size_t)y * src_pitch + 3 * x;
  uint8_t* d = gray + (size_t)y * dst_pitch + x;
  uint8_t o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (x + k < w) {
      const int b = __ldg(s + 3 * k), g = __ldg(s + 3 * k + 1), r = __ldg(s + 3 * k + 2);
      o[k] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
    } else {
      o[k] = 0;
    }
  }
  if (x + 4 <= w && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
    *reinterpret_cast<uchar4*>(d) = make_uchar4(o[0], o[1], o[2], o[3]);
  } else {
    for (int k = 0; k < 4 && x + k < w; k++) d[k] = o[k];
  }
}

int bgr2gray_launch(vo_ctx* c, const uint8_t* d_bgr, int src_pitch, uint8_t* d_gray, int dst_pitch) {
  return bgr2gray_launch_wh(c, d_bgr, src_pitch, d_gray, dst_pitch, c->p.width, c->p.height);
}

int bgr2gray_launch_wh(vo_ctx* c, const uint8_t* d_bgr, int src_pitch, uint8_t* d_gray, int dst_pitch, int w, int h) {
  dim3 b(128), g(div_up(div_up(w, 4), 128), h);
  {
    LaunchScope ls(c, VO_K_PYRAMID);
    bgr2gray_kernel<<<g, b, 0, c->stream>>>(d_bgr, w, h, src_pitch, d_gray, dst_pitch);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ------------------------------------------------------------------------------------
int pyr_alloc(vo_ctx* c, Pyramid& p) {
  int w = c->p.width, h = c->p.height;
  const int cn = c->p.channels;
  p.nlevels = 0;
  p.cn = cn;
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
    L.plane = ((size_t)h + 2 * PAD_Y) * L.pitch;
    const size_t guard = (size_t)GUARD_ROWS * L.pitch;
    VO_CUDA(cudaMalloc(&L.img_alloc, cn * L.plane + 2 * guard));
    L.img = L.img_alloc + guard;
    VO_CUDA(cudaMalloc(&L.deriv, cn * L.plane * sizeof(short2)));
    VO_CUDA(cudaMemsetAsync(L.img_alloc, 0, cn * L.plane + 2 * guard, c->stream));
    VO_CUDA(cudaMemsetAsync(L.deriv, 0, cn * L.plane * sizeof(short2), c->stream));
    p.nlevels = l + 1;
  }
  p.has_deriv = false;
  p.stamp = 0;
  VO_CUDA(cudaMalloc(&p.d_mono, sizeof(int)));
  VO_CUDA(cudaMemsetAsync(p.d_mono, 0, sizeof(int), c->stream));
  // TMA descriptors of the level-0 interiors (the driver entry point is resolved through the
  // runtime, so the library does not link libcuda)
  {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    VO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available in this driver");
      return VO_ERR_CUDA;
    }
    PyrLevel& L0 = p.lv[0];
    const cuuint64_t gdim[2] = {(cuuint64_t)L0.w, (cuuint64_t)L0.h};
    const cuuint64_t gstride[1] = {(cuuint64_t)L0.pitch};
    const cuuint32_t box[2] = {(cuuint32_t)K1_BOXW, (cuuint32_t)K1_BOXH};
    const cuuint32_t estr[2] = {1, 1};
    static_assert(sizeof(CUtensorMap) == sizeof(p.tmap0[0]), "CUtensorMap size");
    for (int k = 0; k < cn; k++) {
      CUtensorMap tm;
      const CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                                        L0.img + k * L0.plane + (size_t)PAD_Y * L0.pitch + PAD_L, gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for %d x %d, pitch %d", (int)r, L0.w, L0.h, L0.pitch);
        return VO_ERR_CUDA;
      }
      memcpy(p.tmap0[k], &tm, sizeof(tm));
    }
  }
  return VO_OK;
}

void pyr_free(Pyramid& p) {
  for (int l = 0; l < p.nlevels; l++) {
    cudaFree(p.lv[l].img_alloc);
    cudaFree(p.lv[l].deriv);
    p.lv[l].img = nullptr;
    p.lv[l].img_alloc = nullptr;
    p.lv[l].deriv = nullptr;
  }
  p.nlevels = 0;
  cudaFree(p.d_mono);
  p.d_mono = nullptr;
}

PyrView pyr_view(const Pyramid& p) {
  PyrView v;
  v.nlevels = p.nlevels;
  v.mono = p.d_mono;
  for (int l = 0; l < MAX_LEVELS; l++) {
    if (l < p.nlevels) {
      v.lv[l].img = p.lv[l].img;
      v.lv[l].deriv = p.lv[l].deriv;
      v.lv[l].w = p.lv[l].w;
      v.lv[l].h = p.lv[l].h;
      v.lv[l].pitch = p.lv[l].pitch;
      v.lv[l].plane = (unsigned)p.lv[l].plane;
    } else {
      v.lv[l] = PyrLevelView{nullptr, nullptr, 0, 0, 0, 0};
    }
  }
  return v;
}

static int scharr_all(vo_ctx* c, Pyramid& p) {
  for (int l = 0; l < p.nlevels; l++) {
    PyrLevel& L = p.lv[l];
    dim3 b(32, 8), g(div_up(L.w + PAD_L + PAD_R, 32), div_up(L.h + 2 * PAD_Y, 8));
    for (int k = 0; k < p.cn; k++) {
      LaunchScope ls(c, VO_K_PYRAMID);
      scharr_kernel<<<g, b, 0, c->stream>>>(L.img + k * L.plane, L.deriv + k * L.plane, L.w, L.h, L.pitch);
    }
  }
  VO_CUDA(cudaGetLastError());
  p.has_deriv = true;
  return VO_OK;
}

// Level 0 interior(s) must already be in place (memcpy2D straight into the padded buffer, or
// pyr_split_bgr for 3-channel images).
int pyr_build(vo_ctx* c, int slot, const uint8_t* d_tight, bool with_deriv) {
  Pyramid& p = c->pyr[slot];
  PyrLevel& L0 = p.lv[0];
  if (d_tight) {
    VO_CUDA(cudaMemcpy2DAsync(L0.img + (size_t)PAD_Y * L0.pitch + PAD_L, L0.pitch, d_tight, L0.w, L0.w, L0.h,
                              cudaMemcpyDeviceToDevice, c->stream));
  }
  K1Args a;
  a.nlevels = p.nlevels;
  a.with_deriv = with_deriv ? 1 : 0;
  for (int l = 0; l < MAX_LEVELS; l++) {
    if (l < p.nlevels) {
      a.lv[l] = K1Level{p.lv[l].img, p.lv[l].deriv, p.lv[l].w, p.lv[l].h, p.lv[l].pitch};
      a.plane[l] = p.lv[l].plane;
    } else {
      a.lv[l] = K1Level{nullptr, nullptr, 0, 0, 0};
      a.plane[l] = 0;
    }
  }
  K1Maps maps;
  for (int k = 0; k < MAX_CN; k++) memcpy(&maps.m[k], p.tmap0[k < p.cn ? k : 0], sizeof(CUtensorMap));
  {
    dim3 g(div_up(L0.w, K1_T0W), div_up(L0.h, K1_T0H), p.cn);
    LaunchScope ls(c, VO_K_PYRAMID);
    pyr_fused_kernel<<<g, K1_THREADS, 0, c->stream>>>(maps, a);
  }
  VO_CUDA(cudaGetLastError());
  p.has_deriv = with_deriv;
  p.stamp = ++c->stamp_counter;
  return VO_OK;
}

// 3-channel input: de-interleave the tight BGR image in d_bgr into the level-0 planes of `slot`.
int pyr_split_bgr(vo_ctx* c, int slot, const uint8_t* d_bgr) {
  PyrLevel& L0 = c->pyr[slot].lv[0];
  dim3 b(128), g(div_up(div_up(L0.w, 4), 128), L0.h);
  {
    LaunchScope ls(c, VO_K_PYRAMID);
    set_int_kernel<<<1, 1, 0, c->stream>>>(c->pyr[slot].d_mono, 1);
    split_planes_kernel<<<g, b, 0, c->stream>>>(d_bgr, L0.w, L0.h, L0.img, L0.pitch, L0.plane, c->pyr[slot].d_mono);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int pyr_ensure_deriv(vo_ctx* c, int slot) {
  Pyramid& p = c->pyr[slot];
  if (p.has_deriv) return VO_OK;
  return scharr_all(c, p);
}

}  // namespace vo
