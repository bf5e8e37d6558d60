// sgbm.cu -- dense stereo: StereoProcess::stereoMatch (reference src/StereoCV.cpp:21-62) =
// cvtColor(BGR2GRAY) x2 + StereoSGBM::create(1, 96, 7, 24, 96, 0, 60, 0, 3000, 5)->compute, and
// StereoProcess::reprojectDisparity (src/StereoCV.cpp:221-250) = reprojectImageTo3D + depth gate.
// SURVEY.md section 8 rows a-11 / (f)-4.  All integer; bit-identical to cv2 4.13.0 (oracle/sgbm.py is the
// stage-by-stage restatement these kernels are compared with).
//
// Stages and HBM layout (W1 = number of columns that get a disparity = width - maxD for minD >= 0, D = numDisparities):
//   K_pre    x-Sobel prefilter + the Birchfield-Tomasi half-pixel min/max of both planes (filtered, raw) of both
//            images -> uchar4 (v, lo, hi, -) per pixel and plane                         [2][H][W] uchar4 x 2
//   K_hsum   a CTA owns a strip of one row: BT pixel costs of the strip + halo for all D are computed into shared
//            memory (the row-wise search over shared-memory strips of the right image), summed over the block
//            width -> hsum[y][x][d] u16                                                  H*W1*D*2 bytes
//   K_vsum   running sum over the block height (rows replicated at the border) -> C[y][x][d] s16
//   K_path   one WARP per path, the D disparities of a pixel spread over the lanes (4 or 8 per lane, one 8/16-byte
//            access per step), predecessor costs in registers as u16x2 pairs, the recurrence in VIADDMNMX.S16x2 /
//            VIMNMX.S16x2 (two disparities per instruction), neighbours d-1/d+1 by two shuffles, min over d by one
//            REDUX; C of the next 16 steps is in flight through a per-warp cp.async ring in shared memory, so the only
//            latency on the chain is the recurrence itself.  A path is one warp and the paths are few (2h horizontal,
//            2(W1+h-1) diagonal, W1 vertical), so all five directions run at once on three streams, each into its own
//            volume: L - C (it lies in [0, P2]) as one byte per cost when P2 <= 255, L as s16 otherwise.  The diagonal
//            and vertical launches are issue-bound and put two paths into a warp (K_path2, D <= 128).
//   K_wta    warp per pixel: S = min(32767, sum of the five L) (every L is >= 0, so OpenCV's two saturating adds
//            collapse into this and the order of the directions is free), winner-take-all, uniqueness test, the
//            neighbours for the sub-pixel step and the right-view disparity (one atomicMin on a packed
//            (cost, 65535 - x) key replaces OpenCV's descending-x first-come rule).
//   K_lr     left-right check, K_median 3x3, K_speckle: union-find connected components over |difference| <=
//            16*speckleRange edges, components of <= speckleWindowSize pixels are cleared (cv::filterSpeckles).
//   K_reproj reprojectImageTo3D on the float-converted 16x disparity + the reference's gate 0.01 < z <= 5 and y flip,
//            then an order-preserving compaction (CUB DeviceSelect) -> points + their pixel indices for the
//            caller's colour lookup.
// A repeated (size, parameters) call replays the whole pipeline as one CUDA graph (sgbm_run); pageable caller buffers
// travel through pinned staging owned by the context.
#include <cub/cub.cuh>

#include "common.cuh"

namespace vo {

constexpr int SG_MAX_COST = 32767;
constexpr int SG_TX = 64;          // columns per K_hsum CTA
constexpr int SG_MAX_D = 256;
constexpr int SG_MAX_R = 5;        // block size <= 11
constexpr unsigned SG_KEY_INIT = 0xFFFFFFFFu;

struct Sgbm {
  int w = 0, h = 0;                         // the per-pixel buffers hold w*h pixels, the volumes cost_elems costs
  size_t cost_elems = 0;
  uint8_t* img[2] = {nullptr, nullptr};     // tight gray images
  uint8_t* bgr[2] = {nullptr, nullptr};     // tight BGR staging (vo_stereo_match)
  uchar4* pl = nullptr;                     // [2 images][2 planes][h][w]
  uint16_t* hsum = nullptr;                 // also reused as S2
  int16_t* C = nullptr;
  int16_t* S = nullptr;
  int16_t* vol[3] = {nullptr, nullptr, nullptr};   // path-cost volumes L4, L1, L3 (L0 reuses hsum, L2 is S)
  cudaStream_t stream2 = nullptr, stream3 = nullptr;   // horizontal / diagonal paths run beside the vertical ones
  cudaEvent_t ev_h = nullptr, ev_d = nullptr, ev_fork = nullptr;
  unsigned* key2 = nullptr;                 // right-view (cost, x) keys
  uint2* rec = nullptr;                     // winner records
  int16_t* disp[3] = {nullptr, nullptr, nullptr};   // raw WTA, after LR check, after median (+ speckle in place)
  int* label = nullptr;
  int* count = nullptr;
  float3* xyz = nullptr;                    // reprojection, per pixel
  uint8_t* keep = nullptr;
  float3* xyz_out = nullptr;
  int* idx_out = nullptr;
  int* d_n = nullptr;
  double* dQ = nullptr;
  uint8_t* h_in[2] = {nullptr, nullptr};    // pinned staging for pageable caller memory
  uint8_t* h_out = nullptr;
  size_t h_in_bytes = 0, h_out_bytes = 0;
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  bool have_disp = false;
  cudaGraphExec_t gexec = nullptr;          // the captured pipeline for (g_w, g_h, g_params)
  int g_w = 0, g_h = 0, g_launches = 0;
  vo_sgbm_params g_params = {};
  bool graph_failed = false;
  bool staged = false;                      // the last call recorded per-stage events
  float pipeline_ms = 0.f;
  bool stage_events = true;                 // ev[2..7] are recorded (plain launches); false while capturing / replaying
  int last_w = 0, last_h = 0;
  cudaEvent_t ev[10] = {nullptr};
  float ms[9] = {0};
};

void sgbm_free(vo_ctx* c) {
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  if (!s) return;
  void* dev[] = {s->img[0], s->img[1], s->bgr[0], s->bgr[1], s->pl, s->hsum, s->C, s->S, s->vol[0], s->vol[1], s->vol[2], s->key2, s->rec, s->disp[0], s->disp[1],
                 s->disp[2], s->label, s->count, s->xyz, s->keep, s->xyz_out, s->idx_out, s->d_n, s->dQ, s->cub_tmp};
  for (void* p : dev) cudaFree(p);
  cudaFreeHost(s->h_in[0]);
  cudaFreeHost(s->h_in[1]);
  cudaFreeHost(s->h_out);
  for (auto e : s->ev)
    if (e) cudaEventDestroy(e);
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  if (s->ev_h) cudaEventDestroy(s->ev_h);
  if (s->ev_d) cudaEventDestroy(s->ev_d);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->stream2) cudaStreamDestroy(s->stream2);
  if (s->stream3) cudaStreamDestroy(s->stream3);
  delete s;
  c->sgbm = nullptr;
}

static void sgbm_release_buffers(Sgbm* s) {
  void** dev[] = {(void**)&s->img[0], (void**)&s->img[1], (void**)&s->bgr[0], (void**)&s->bgr[1], (void**)&s->pl,
                  (void**)&s->hsum, (void**)&s->C, (void**)&s->S, (void**)&s->vol[0], (void**)&s->vol[1], (void**)&s->vol[2],
                  (void**)&s->key2, (void**)&s->rec, (void**)&s->disp[0],
                  (void**)&s->disp[1], (void**)&s->disp[2], (void**)&s->label, (void**)&s->count, (void**)&s->xyz,
                  (void**)&s->keep, (void**)&s->xyz_out, (void**)&s->idx_out, (void**)&s->cub_tmp};
  for (void** p : dev) {
    cudaFree(*p);
    *p = nullptr;
  }
}

// ------------------------------------------------------------------------------------ K_pre
// calcPixelCostBT, first half: the clipped x-Sobel plane and the raw plane of one image row; OpenCV writes tab[0]
// (= ftzero) into both row ends of both planes.
__device__ __forceinline__ int sg_plane_value(const uint8_t* __restrict__ r0, const uint8_t* __restrict__ rn,
                                              const uint8_t* __restrict__ rs, int x, int w, int ftzero, int plane) {
  if (x <= 0 || x >= w - 1) return ftzero;
  if (plane) return r0[x];
  int s = ((int)r0[x + 1] - (int)r0[x - 1]) * 2 + (int)rn[x + 1] - (int)rn[x - 1] + (int)rs[x + 1] - (int)rs[x - 1];
  s = min(max(s, -ftzero), ftzero);
  return s + ftzero;
}

__global__ void sgbm_prefilter_kernel(const uint8_t* __restrict__ imgL, const uint8_t* __restrict__ imgR, int w, int h,
                                      int ftzero, uchar4* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int im = blockIdx.z;
  if (x >= w) return;
  const uint8_t* img = im ? imgR : imgL;
  const uint8_t* r0 = img + (size_t)y * w;
  const uint8_t* rn = y > 0 ? r0 - w : r0;
  const uint8_t* rs = y < h - 1 ? r0 + w : r0;
#pragma unroll
  for (int plane = 0; plane < 2; plane++) {
    const int v = sg_plane_value(r0, rn, rs, x, w, ftzero, plane);
    const int vl = x > 0 ? (v + sg_plane_value(r0, rn, rs, x - 1, w, ftzero, plane)) / 2 : v;
    const int vr = x < w - 1 ? (v + sg_plane_value(r0, rn, rs, x + 1, w, ftzero, plane)) / 2 : v;
    const int lo = min(min(vl, vr), v), hi = max(max(vl, vr), v);
    out[(((size_t)im * 2 + plane) * h + y) * w + x] = make_uchar4((unsigned char)v, (unsigned char)lo, (unsigned char)hi, 0);
  }
}

// ------------------------------------------------------------------------------------ K_hsum
// pixel cost of calcPixelCostBT for (x1 = column in the left image, d): both planes, the raw one >> 2
__device__ __forceinline__ int sg_bt(uchar4 u, uchar4 v, int shift) {
  const int c0 = max(max(0, (int)u.x - (int)v.z), (int)v.y - (int)u.x);
  const int c1 = max(max(0, (int)v.x - (int)u.z), (int)u.y - (int)v.x);
  return min(c0, c1) >> shift;
}

__global__ void __launch_bounds__(256, 4)
sgbm_hsum_kernel(const uchar4* __restrict__ pl, int w, int h, int W1, int D, int minD, int minX1, int R,
                 uint16_t* __restrict__ hsum) {
  extern __shared__ __align__(16) unsigned char sg_smem[];
  const int y = blockIdx.y;
  const int x0 = blockIdx.x * SG_TX;
  const int ncol = SG_TX + 2 * R;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint16_t* sg_pix = reinterpret_cast<uint16_t*>(sg_smem);                       // [ncol][D] pixel costs
  uchar4* sg_right = reinterpret_cast<uchar4*>(sg_smem + (size_t)ncol * D * 2);  // [2 planes][ncol + D - 1] right-image strip
  const uchar4* L0 = pl + ((size_t)0 * h + y) * w;   // left filtered
  const uchar4* L1 = pl + ((size_t)1 * h + y) * w;   // left raw
  const uchar4* R0 = pl + ((size_t)2 * h + y) * w;   // right filtered
  const uchar4* R1 = pl + ((size_t)3 * h + y) * w;   // right raw
  // the strip of the right image this CTA searches: x2 = x1 - d for every column of the strip (+ halo) and every d
  const int x1_lo = min(max(x0 - R, 0), W1 - 1) + minX1;
  const int x2_base = x1_lo - minD - (D - 1);
  const int nright = ncol + D - 1;
  for (int i = threadIdx.x; i < nright; i += blockDim.x) {
    const int x2 = min(x2_base + i, w - 1);          // entries beyond the last column are never used
    sg_right[i] = __ldg(R0 + x2);
    sg_right[nright + i] = __ldg(R1 + x2);
  }
  __syncthreads();
  // pixel costs of the strip + halo: a warp per column, the lanes over the disparities
  for (int xi = wid; xi < ncol; xi += 8) {
    const int xw = min(max(x0 - R + xi, 0), W1 - 1);   // columns replicate at the border of the computed range
    const int x1 = xw + minX1;
    const uchar4 u0 = __ldg(L0 + x1), u1 = __ldg(L1 + x1);
    const uchar4* r0 = sg_right + (x1 - minD - x2_base);
    const uchar4* r1 = r0 + nright;
    uint16_t* dst = sg_pix + xi * D;
    for (int dd = lane; dd < D; dd += 32) dst[dd] = (uint16_t)(sg_bt(u0, r0[-dd], 0) + sg_bt(u1, r1[-dd], 2));
  }
  __syncthreads();
  // box width: a warp owns SG_TX / 8 consecutive columns and slides the sum along them
  constexpr int CPW = SG_TX / 8;
  const int nb = 2 * R + 1;
  const int xo = wid * CPW;
  const int Dw = D >> 1;                                // two disparities per 32-bit word
  for (int pp = lane; pp < Dw; pp += 32) {
    const unsigned* px = reinterpret_cast<const unsigned*>(sg_pix) + xo * Dw + pp;
    unsigned* hp = reinterpret_cast<unsigned*>(hsum + ((size_t)y * W1 + x0 + xo) * D) + pp;
    unsigned sum = 0;
    for (int k = 0; k < nb; k++) sum += px[k * Dw];
#pragma unroll
    for (int cI = 0; cI < CPW; cI++) {
      if (x0 + xo + cI < W1) *hp = sum;
      hp += Dw;
      if (cI + 1 < CPW) sum = sum + px[(cI + nb) * Dw] - px[cI * Dw];
    }
  }
}

// ------------------------------------------------------------------------------------ K_vsum
// Eight costs per thread as four u16x2 words, added and subtracted as plain 32-bit integers: a window sum is at most
// 32767 and never smaller than the row that leaves it, so no carry or borrow crosses the halves.
__device__ __forceinline__ uint4 sg_add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint4 sg_sub4(uint4 a, uint4 b) { return make_uint4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

__global__ void __launch_bounds__(256)
sgbm_vsum_kernel(const uint4* __restrict__ hsum, int h, size_t row_vecs, int R, int rows_per_strip, uint4* __restrict__ C) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= row_vecs) return;
  const int ya = blockIdx.y * rows_per_strip;
  const int yb = min(ya + rows_per_strip, h);
  if (ya >= yb) return;
  uint4 s = make_uint4(0u, 0u, 0u, 0u);
  for (int k = -R; k <= R; k++) s = sg_add4(s, __ldg(hsum + (size_t)min(max(ya + k, 0), h - 1) * row_vecs + e));
#pragma unroll 4
  for (int y = ya; y < yb; y++) {
    C[(size_t)y * row_vecs + e] = s;
    const uint4 in = __ldg(hsum + (size_t)min(y + R + 1, h - 1) * row_vecs + e);
    const uint4 out = __ldg(hsum + (size_t)max(y - R, 0) * row_vecs + e);
    s = sg_sub4(sg_add4(s, in), out);
  }
}

// ------------------------------------------------------------------------------------ K_path
template <int DPL> struct SgVec;
template <> struct SgVec<4> { using T = uint2; static constexpr int STAGES = 16; };
template <> struct SgVec<8> { using T = uint4; static constexpr int STAGES = 8; };
constexpr int SG_WPB = 4;          // warps (paths) per CTA of the path kernel

// costs are non-negative int16: zero-extension is enough
template <int DPL>
__device__ __forceinline__ void sg_unpack(const typename SgVec<DPL>::T& v, int* o) {
  const unsigned* p = reinterpret_cast<const unsigned*>(&v);
#pragma unroll
  for (int k = 0; k < DPL / 2; k++) {
    o[2 * k] = (int)(p[k] & 0xFFFFu);
    o[2 * k + 1] = (int)(p[k] >> 16);
  }
}
template <int DPL>
__device__ __forceinline__ typename SgVec<DPL>::T sg_pack(const int* o) {
  typename SgVec<DPL>::T v;
  unsigned* p = reinterpret_cast<unsigned*>(&v);
#pragma unroll
  for (int k = 0; k < DPL / 2; k++) p[k] = __byte_perm((unsigned)o[2 * k], (unsigned)o[2 * k + 1], 0x5410);
  return v;
}

template <int BYTES>
__device__ __forceinline__ void sg_cp_async(unsigned dst, const void* src) {
  if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sg_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sg_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct SgWta {
  uint2* rec;           // [h][w] winner records (d | minS << 16, S[d-1] | S[d+1] << 16); x = 0xFFFFFFFF: no disparity
  unsigned* key2;       // [h][w] right-view (cost << 16 | 65535 - x) keys, pre-filled with SG_KEY_INIT
  int w, minD, minX1, uniq;
};

// One warp per path: out(p, d) = L_r(p, d).  Directions: 0 left-to-right, 4 right-to-left (paths = rows), 1 down-right,
// 2 down, 3 down-left; blockIdx.y selects (dir_a -> out_a) or (dir_b -> out_b), so two directions share a launch.
// C of the steps ahead travels through a per-warp shared-memory ring filled by cp.async: completion is tracked per
// commit group, in order, so a step never waits for a copy younger than its own (register prefetch could not do
// that: the scoreboards a load waits on are shared with the younger loads in flight).
// NARROW (P2 <= 255): the volume holds L - C, which lies in [0, P2], as one byte per disparity -- half the traffic.
template <int DPL, bool NARROW>
__global__ void __launch_bounds__(SG_WPB * 32)
sgbm_path_kernel(const int16_t* __restrict__ C, void* __restrict__ out_a, void* __restrict__ out_b, int W1, int H,
                 int D, int P1, int P2, int dir_a, int dir_b, int npaths) {
  using V = typename SgVec<DPL>::T;
  constexpr int NST = SgVec<DPL>::STAGES;
  constexpr int VB = 2 * DPL;                 // bytes per lane
  constexpr int STG = 32 * VB;
  extern __shared__ __align__(16) unsigned char sg_ring[];   // [warp][stage][lane][VB]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int p = blockIdx.x * SG_WPB + wib;
  if (p >= npaths) return;
  const int dir = blockIdx.y ? dir_b : dir_a;
  void* out = blockIdx.y ? out_b : out_a;
  int x, y, n, sx, sy;
  switch (dir) {
    case 0: x = 0; y = p; n = W1; sx = 1; sy = 0; break;
    case 4: x = W1 - 1; y = p; n = W1; sx = -1; sy = 0; break;
    case 2: x = p; y = 0; n = H; sx = 0; sy = 1; break;
    case 1:
      if (p < W1) { x = p; y = 0; } else { x = 0; y = p - W1 + 1; }
      n = min(H - y, W1 - x); sx = 1; sy = 1; break;
    default:
      if (p < W1) { x = p; y = 0; } else { x = W1 - 1; y = p - W1 + 1; }
      n = min(H - y, x + 1); sx = -1; sy = 1; break;
  }
  if (dir == 2 && p >= W1) return;             // the vertical direction has fewer paths than its diagonal partner
  const bool active = DPL * lane < D;          // lanes beyond D work on lane 0's data and never store
  const bool last = DPL * (lane + 1) >= D;
  const long long step = ((long long)sy * W1 + sx) * D;       // elements per step along the path
  const size_t base = ((size_t)y * W1 + x) * D + (active ? DPL * lane : 0);
  const int16_t* gC = C + base;                // the step the next cp.async fetches
  constexpr int OB = NARROW ? 1 : 2;           // bytes per stored cost
  unsigned char* wS = reinterpret_cast<unsigned char*>(out) + base * OB;   // the current step
  unsigned char* ring = sg_ring + (size_t)wib * NST * STG + lane * VB;
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);

#pragma unroll
  for (int k = 0; k < NST; k++) {
    if (k < n) {
      sg_cp_async<VB>(ring_s + k * STG, gC);
      gC += step;
    }
    sg_cp_commit();
  }
  // Two disparities per register (s16x2) and per instruction: VIADDMNMX.S16x2 / VIMNMX.S16x2 (the DPX instructions).
  // Lp[q] = (L(2q), L(2q + 1)).  The value standing in for the missing neighbours d = -1 and d = D only has to
  // satisfy PAD + P1 >= min_k L + P2; 32767 - P1 does (C + P2 <= 32767 is checked on the host) and cannot wrap.
  constexpr int NW = DPL / 2;
  unsigned Lp[NW];
#pragma unroll
  for (int q = 0; q < NW; q++) Lp[q] = 0u;
  int minp = 0;
  const unsigned PAD2 = (unsigned)(SG_MAX_COST - P1) * 0x10001u;
  const unsigned P1x2 = (unsigned)P1 * 0x10001u;

  for (int i0 = 0; i0 < n; i0 += NST) {
#pragma unroll
    for (int k = 0; k < NST; k++) {
      const int i = i0 + k;
      if (i >= n) break;
      sg_cp_wait<NST - 1>();                  // the group of step i (and every older one) has landed
      const V cvec = *reinterpret_cast<const V*>(ring + k * STG);
      const unsigned* cw = reinterpret_cast<const unsigned*>(&cvec);
      if (i + NST < n) {                      // refill the stage that was just read
        sg_cp_async<VB>(ring_s + k * STG, gC);
        gC += step;
      }
      sg_cp_commit();
      // formula 13 of the SGM paper as OpenCV evaluates it
      unsigned left = __shfl_up_sync(0xffffffffu, Lp[NW - 1], 1);     // its high half is L(d - 1) of this lane's first d
      unsigned right = __shfl_down_sync(0xffffffffu, Lp[0], 1);       // its low half is L(d + 1) of this lane's last d
      if (lane == 0) left = PAD2;
      if (last) right = PAD2;
      const unsigned minp2 = (unsigned)minp * 0x10001u;
      const unsigned delta2 = minp2 + (unsigned)P2 * 0x10001u;
      unsigned Ln[NW], rel[NW];
#pragma unroll
      for (int q = 0; q < NW; q++) {
        const unsigned lo = __byte_perm(q ? Lp[q - 1] : left, Lp[q], 0x5432);             // (L(2q - 1), L(2q))
        const unsigned hi = __byte_perm(Lp[q], q < NW - 1 ? Lp[q + 1] : right, 0x5432);   // (L(2q + 1), L(2q + 2))
        unsigned t = __viaddmin_s16x2(lo, P1x2, Lp[q]);
        t = __viaddmin_s16x2(hi, P1x2, t);
        t = __vmins2(t, delta2);
        rel[q] = t - minp2;                   // L - C, in [0, P2] per half: no borrow between the halves
        Ln[q] = rel[q] + cw[q];               // <= 32767 per half: no carry
      }
      unsigned mw = Ln[0];
#pragma unroll
      for (int q = 1; q < NW; q++) mw = __vmins2(mw, Ln[q]);
      int m = min((int)(mw & 0xFFFFu), (int)(mw >> 16));
      if (!active) m = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < NW; q++) Lp[q] = Ln[q];
      minp = __reduce_min_sync(0xffffffffu, m);
      if (active) {
        if (NARROW) {
          if (DPL == 4) *reinterpret_cast<unsigned*>(wS) = __byte_perm(rel[0], rel[1], 0x6420);
          else *reinterpret_cast<uint2*>(wS) = make_uint2(__byte_perm(rel[0], rel[1], 0x6420),
                                                           __byte_perm(rel[NW - 2], rel[NW - 1], 0x6420));
        } else {
          V o;
          unsigned* ow = reinterpret_cast<unsigned*>(&o);
#pragma unroll
          for (int q = 0; q < NW; q++) ow[q] = Ln[q];
          *reinterpret_cast<V*>(wS) = o;
        }
      }
      wS += step * OB;
    }
  }
}

// Two paths per warp (D <= 128): each half-warp walks its own path, 8 disparities per lane.  The diagonal and the
// vertical directions have thousands of paths and are bound by instruction issue, not by the latency of one path, so
// the fixed per-step overhead (ring, pointers, loop) is shared by two paths; the price is the min over d as four
// shuffle steps inside the half-warp instead of one REDUX.  (The horizontal pair stays on the one-path kernel: there
// the latency of a path is the whole launch.)
template <bool NARROW>
__global__ void __launch_bounds__(SG_WPB * 32)
sgbm_path2_kernel(const int16_t* __restrict__ C, void* __restrict__ out_a, void* __restrict__ out_b, int W1, int H, int D,
                  int P1, int P2, int dir_a, int dir_b, int npaths) {
  constexpr int NST = 8, NW = 4, STG = 32 * 16;
  extern __shared__ __align__(16) unsigned char sg_ring[];   // [warp][stage][lane][16]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int hl = lane & 15;
  const int p = 2 * (blockIdx.x * SG_WPB + wib) + (lane >> 4);
  const int dir = blockIdx.y ? dir_b : dir_a;
  void* out = blockIdx.y ? out_b : out_a;
  int x = 0, y = 0, n = 0, sx = 0, sy = 1;
  if (p < npaths) {
    switch (dir) {
      case 2: x = p; y = 0; n = H; sx = 0; break;
      case 1:
        if (p < W1) { x = p; y = 0; } else { x = 0; y = p - W1 + 1; }
        n = min(H - y, W1 - x); sx = 1; break;
      default:
        if (p < W1) { x = p; y = 0; } else { x = W1 - 1; y = p - W1 + 1; }
        n = min(H - y, x + 1); sx = -1; break;
    }
  }
  const int nmax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
  if (nmax == 0) return;
  const bool active = 8 * hl < D;
  const bool last = 8 * (hl + 1) >= D;
  const long long step = ((long long)sy * W1 + sx) * D;
  const size_t base = ((size_t)y * W1 + x) * D + (active ? 8 * hl : 0);
  const int16_t* gC = C + base;
  constexpr int OB = NARROW ? 1 : 2;
  unsigned char* wS = reinterpret_cast<unsigned char*>(out) + base * OB;
  unsigned char* ring = sg_ring + (size_t)wib * NST * STG + lane * 16;
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
  const bool loads = active && n > 0;

#pragma unroll
  for (int k = 0; k < NST; k++) {
    if (loads && k < n) {
      sg_cp_async<16>(ring_s + k * STG, gC);
      gC += step;
    }
    sg_cp_commit();
  }
  unsigned Lp[NW];
#pragma unroll
  for (int q = 0; q < NW; q++) Lp[q] = 0u;
  int minp = 0;
  const unsigned PAD2 = (unsigned)(SG_MAX_COST - P1) * 0x10001u;
  const unsigned P1x2 = (unsigned)P1 * 0x10001u;

  for (int i0 = 0; i0 < nmax; i0 += NST) {
#pragma unroll
    for (int k = 0; k < NST; k++) {
      const int i = i0 + k;
      if (i >= nmax) break;
      sg_cp_wait<NST - 1>();
      const uint4 cvec = *reinterpret_cast<const uint4*>(ring + k * STG);
      const unsigned cw[NW] = {cvec.x, cvec.y, cvec.z, cvec.w};
      if (loads && i + NST < n) {
        sg_cp_async<16>(ring_s + k * STG, gC);
        gC += step;
      }
      sg_cp_commit();
      unsigned left = __shfl_up_sync(0xffffffffu, Lp[NW - 1], 1, 16);
      unsigned right = __shfl_down_sync(0xffffffffu, Lp[0], 1, 16);
      if (hl == 0) left = PAD2;
      if (last) right = PAD2;
      const unsigned minp2 = (unsigned)minp * 0x10001u;
      const unsigned delta2 = minp2 + (unsigned)P2 * 0x10001u;
      unsigned Ln[NW], rel[NW];
#pragma unroll
      for (int q = 0; q < NW; q++) {
        const unsigned lo = __byte_perm(q ? Lp[q - 1] : left, Lp[q], 0x5432);
        const unsigned hi = __byte_perm(Lp[q], q < NW - 1 ? Lp[q + 1] : right, 0x5432);
        unsigned t = __viaddmin_s16x2(lo, P1x2, Lp[q]);
        t = __viaddmin_s16x2(hi, P1x2, t);
        t = __vmins2(t, delta2);
        rel[q] = t - minp2;
        Ln[q] = rel[q] + cw[q];
      }
      unsigned mw = __vmins2(__vmins2(Ln[0], Ln[1]), __vmins2(Ln[2], Ln[3]));
      int m = min((int)(mw & 0xFFFFu), (int)(mw >> 16));
      if (!active) m = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < NW; q++) Lp[q] = Ln[q];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o, 16));
      minp = m;
      if (active && i < n) {
        if (NARROW) *reinterpret_cast<uint2*>(wS) = make_uint2(__byte_perm(rel[0], rel[1], 0x6420), __byte_perm(rel[2], rel[3], 0x6420));
        else *reinterpret_cast<uint4*>(wS) = make_uint4(Ln[0], Ln[1], Ln[2], Ln[3]);
      }
      wS += step * OB;
    }
  }
}

// ------------------------------------------------------------------------------------ K_wta
// One warp per pixel (grid-stride): S = min(32767, L0 + L1 + L2 + L3 + L4) -- every L is >= 0, so OpenCV's two
// saturating adds collapse into this -- then winner-take-all (first minimum), uniqueness test, the neighbours of the
// minimum for the sub-pixel step, and the right-view disparity (atomicMin on (cost << 16 | 65535 - x)).
constexpr int SG_WTA_PIX = 8;      // consecutive pixels of a row per warp of the WTA kernel

template <int DPL, bool NARROW>
struct SgWtaIn {                   // what one lane loads for one pixel: C (narrow only) and the five volumes
  typename SgVec<DPL>::T c;
  unsigned n[5][DPL / 4];          // narrow: 4 costs per word
  typename SgVec<DPL>::T w[5];     // wide: int16 costs
};

template <int DPL, bool NARROW>
__device__ __forceinline__ void sg_wta_load(SgWtaIn<DPL, NARROW>& in, const int16_t* __restrict__ C, const void* const* vol,
                                            size_t off) {
  using V = typename SgVec<DPL>::T;
  if (NARROW) {
    in.c = __ldg(reinterpret_cast<const V*>(C + off));
#pragma unroll
    for (int v = 0; v < 5; v++) {
      const unsigned char* p = reinterpret_cast<const unsigned char*>(vol[v]) + off;
      if (DPL == 4) {
        in.n[v][0] = __ldg(reinterpret_cast<const unsigned*>(p));
      } else {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        in.n[v][0] = t.x;
        in.n[v][DPL / 4 - 1] = t.y;
      }
    }
  } else {
#pragma unroll
    for (int v = 0; v < 5; v++) in.w[v] = __ldg(reinterpret_cast<const V*>(reinterpret_cast<const int16_t*>(vol[v]) + off));
  }
}

// S of one pixel: narrow volumes hold L - C, so S = 5 C + their sum; four bytes are widened to two u16x2 words by PRMT
// and added as plain 32-bit integers (five bytes sum to at most 1275, no carry between the halves)
template <int DPL, bool NARROW>
__device__ __forceinline__ void sg_wta_sum(const SgWtaIn<DPL, NARROW>& in, int* tot) {
  if (NARROW) {
    unsigned acc[DPL / 2];
#pragma unroll
    for (int k = 0; k < DPL / 2; k++) acc[k] = 0u;
#pragma unroll
    for (int v = 0; v < 5; v++) {
#pragma unroll
      for (int q = 0; q < DPL / 4; q++) {
        acc[2 * q] += __byte_perm(in.n[v][q], 0u, 0x4140);
        acc[2 * q + 1] += __byte_perm(in.n[v][q], 0u, 0x4342);
      }
    }
    int cv[DPL];
    sg_unpack<DPL>(in.c, cv);
#pragma unroll
    for (int k = 0; k < DPL / 2; k++) {
      tot[2 * k] = min(5 * cv[2 * k] + (int)(acc[k] & 0xFFFFu), SG_MAX_COST);
      tot[2 * k + 1] = min(5 * cv[2 * k + 1] + (int)(acc[k] >> 16), SG_MAX_COST);
    }
  } else {
    int t[DPL];
#pragma unroll
    for (int j = 0; j < DPL; j++) tot[j] = 0;
#pragma unroll
    for (int v = 0; v < 5; v++) {
      sg_unpack<DPL>(in.w[v], t);
#pragma unroll
      for (int j = 0; j < DPL; j++) tot[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < DPL; j++) tot[j] = min(tot[j], SG_MAX_COST);
  }
}

// A warp owns SG_WTA_PIX consecutive pixels of row blockIdx.y; the loads of the next pixel are issued before the
// current one is reduced.
template <int DPL, bool NARROW>
__global__ void __launch_bounds__(256)
sgbm_wta_kernel(const int16_t* __restrict__ C, const void* __restrict__ v0, const void* __restrict__ v1,
                const void* __restrict__ v2, const void* __restrict__ v3, const void* __restrict__ v4, int W1, int H, int D,
                SgWta wta) {
  const int lane = threadIdx.x & 31;
  const bool active = DPL * lane < D;
  const int py = blockIdx.y;
  const int xa = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SG_WTA_PIX;
  if (xa >= W1) return;
  const int xb = min(xa + SG_WTA_PIX, W1);
  const void* const vol[5] = {v0, v1, v2, v3, v4};
  size_t off = ((size_t)py * W1 + xa) * D + (active ? DPL * lane : 0);
  SgWtaIn<DPL, NARROW> cur, nxt;
  sg_wta_load<DPL, NARROW>(nxt, C, vol, off);
  unsigned* key2row = wta.key2 + (size_t)py * wta.w;
  uint2* recrow = wta.rec + (size_t)py * wta.w + wta.minX1;
  for (int px = xa; px < xb; px++) {
    cur = nxt;
    off += D;
    if (px + 1 < xb) sg_wta_load<DPL, NARROW>(nxt, C, vol, off);
    int tot[DPL];
    sg_wta_sum<DPL, NARROW>(cur, tot);
    int mykey = (tot[0] << 8) | (DPL * lane);
#pragma unroll
    for (int j = 1; j < DPL; j++) mykey = min(mykey, (tot[j] << 8) | (DPL * lane + j));
    if (!active) mykey = 0x7fffffff;
    const int key = __reduce_min_sync(0xffffffffu, mykey);
    const int e_lo = __shfl_up_sync(0xffffffffu, tot[DPL - 1], 1);     // S[d - 1] of this lane's first d
    const int e_hi = __shfl_down_sync(0xffffffffu, tot[0], 1);         // S[d + 1] of this lane's last d
    const int minS = key >> 8, bd = key & 255;
    bool rej = false;
    if (wta.uniq > 0) {
      bool pr = false;
#pragma unroll
      for (int j = 0; j < DPL; j++)
        pr |= (tot[j] * (100 - wta.uniq) < minS * 100) && (abs(bd - (DPL * lane + j)) > 1);
      rej = __any_sync(0xffffffffu, pr && active);
    }
    if (mykey == key && !rej) {             // the lane that holds the minimum finishes the pixel
      int sm = 0, sp = 0;
#pragma unroll
      for (int j = 0; j < DPL; j++) {
        const bool hit = ((tot[j] << 8) | (DPL * lane + j)) == key;    // exactly one j of this lane
        sm = hit ? (j > 0 ? tot[j - 1] : e_lo) : sm;
        sp = hit ? (j < DPL - 1 ? tot[j + 1] : e_hi) : sp;
      }
      const int x2 = px + wta.minX1 - bd - wta.minD;
      if (minS < SG_MAX_COST) atomicMin(key2row + x2, ((unsigned)minS << 16) | (unsigned)(0xFFFF - px));
      recrow[px] = make_uint2((unsigned)bd | ((unsigned)minS << 16), (unsigned)sm | ((unsigned)sp << 16));
    }
  }
}

// ------------------------------------------------------------------------------------ K_lr, K_median
__global__ void sgbm_fill_kernel(uint2* __restrict__ rec, unsigned* __restrict__ key2, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rec[i] = make_uint2(0xFFFFFFFFu, 0u);
  key2[i] = SG_KEY_INIT;
}

__device__ __forceinline__ int sg_disp2(const unsigned* __restrict__ krow, int xr, int minX1, int inv) {
  const unsigned k = krow[xr];
  if (k == SG_KEY_INIT) return inv;
  const int xw = 0xFFFF - (int)(k & 0xFFFFu);
  return xw + minX1 - xr;     // = d + minD of the winning left pixel
}

// sub-pixel step on the winner record, then the left-right check of computeDisparitySGBM
__global__ void sgbm_lrcheck_kernel(const uint2* __restrict__ rec, const unsigned* __restrict__ key2, int w, int h, int D,
                                    int minD, int minX1, int maxX1, int d12, int16_t* __restrict__ out,
                                    int16_t* __restrict__ out_raw) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const int inv = (minD - 1) * 16;
  int d1 = inv;
  const uint2 r = rec[(size_t)y * w + x];
  if (r.x != 0xFFFFFFFFu) {
    const int bd = (int)(r.x & 0xFFFFu), minS = (int)(r.x >> 16);
    int dv = bd * 16;
    if (0 < bd && bd < D - 1) {
      // parabola through (d-1, S[d-1]), (d, S[d]), (d+1, S[d+1]); C integer division
      const int sm = (int)(r.y & 0xFFFFu), sp = (int)(r.y >> 16);
      const int denom2 = max(sm + sp - 2 * minS, 1);
      dv += ((sm - sp) * 16 + denom2) / (denom2 * 2);
    }
    d1 = dv + minD * 16;
  }
  if (out_raw) out_raw[(size_t)y * w + x] = (int16_t)d1;
  if (x >= minX1 && x < maxX1 && d1 != inv) {
    const unsigned* krow = key2 + (size_t)y * w;
    const int dlo = d1 >> 4, dhi = (d1 + 15) >> 4;
    const int xl = x - dlo, xh = x - dhi;
    bool f1 = false, f2 = false;
    if (0 <= xl && xl < w) {
      const int v = sg_disp2(krow, xl, minX1, inv);
      f1 = v >= minD && abs(v - dlo) > d12;
    }
    if (0 <= xh && xh < w) {
      const int v = sg_disp2(krow, xh, minX1, inv);
      f2 = v >= minD && abs(v - dhi) > d12;
    }
    if (f1 && f2) d1 = inv;
  }
  out[(size_t)y * w + x] = (int16_t)d1;
}

__device__ __forceinline__ void sg_sort2(int& a, int& b) {
  const int t = min(a, b);
  b = max(a, b);
  a = t;
}

__global__ void sgbm_median3_kernel(const int16_t* __restrict__ in, int w, int h, int16_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const int xm = max(x - 1, 0), xp = min(x + 1, w - 1);
  const int16_t* r0 = in + (size_t)max(y - 1, 0) * w;
  const int16_t* r1 = in + (size_t)y * w;
  const int16_t* r2 = in + (size_t)min(y + 1, h - 1) * w;
  int p0 = r0[xm], p1 = r0[x], p2 = r0[xp], p3 = r1[xm], p4 = r1[x], p5 = r1[xp], p6 = r2[xm], p7 = r2[x], p8 = r2[xp];
  // median-of-9 exchange network
  sg_sort2(p1, p2); sg_sort2(p4, p5); sg_sort2(p7, p8); sg_sort2(p0, p1);
  sg_sort2(p3, p4); sg_sort2(p6, p7); sg_sort2(p1, p2); sg_sort2(p4, p5);
  sg_sort2(p7, p8); sg_sort2(p0, p3); sg_sort2(p5, p8); sg_sort2(p4, p7);
  sg_sort2(p3, p6); sg_sort2(p1, p4); sg_sort2(p2, p5); sg_sort2(p4, p7);
  sg_sort2(p4, p2); sg_sort2(p6, p4); sg_sort2(p4, p2);
  out[(size_t)y * w + x] = (int16_t)p4;
}

// ------------------------------------------------------------------------------------ K_speckle (cv::filterSpeckles)
// labels only ever decrease and always name an ancestor, so shortening a chain (path halving) is safe under
// concurrent unions
__device__ __forceinline__ int sg_find(volatile int* lab, int i) {
  int p = lab[i];
  while (p != i) {
    const int g = lab[p];
    if (g != p) lab[i] = g;
    i = p;
    p = g;
  }
  return i;
}

__device__ __forceinline__ void sg_union(int* lab, int a, int b) {
  for (;;) {
    a = sg_find(lab, a);
    b = sg_find(lab, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(lab + b, a);     // hook the larger root under the smaller one
    if (old == b) return;
    b = old;
  }
}

// Row pass: a CTA per image row labels every valid pixel with the flat index of the first pixel of its horizontal
// run (inclusive max-scan of the run starts); invalid pixels get -1.  Only run starts ever become tree nodes.
__global__ void __launch_bounds__(256)
sgbm_cc_rows_kernel(const int16_t* __restrict__ d, int w, int new_val, int max_diff, int* __restrict__ lab,
                    int* __restrict__ cnt) {
  using Scan = cub::BlockScan<int, 256>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int carry;
  const int y = blockIdx.x;
  const int16_t* row = d + (size_t)y * w;
  if (threadIdx.x == 0) carry = -1;
  __syncthreads();
  for (int x0 = 0; x0 < w; x0 += 256) {
    const int x = x0 + threadIdx.x;
    int v = new_val, f = -1;
    if (x < w) {
      v = row[x];
      if (v != new_val) {
        bool conn = false;
        if (x > 0) {
          const int l = row[x - 1];
          conn = l != new_val && abs(v - l) <= max_diff;
        }
        f = conn ? -1 : x;
      }
    }
    int r;
    Scan(tmp).InclusiveScan(f, r, cub::Max());
    r = max(r, carry);
    __syncthreads();
    if (threadIdx.x == 255) carry = r;
    if (x < w) {
      lab[(size_t)y * w + x] = v != new_val ? y * w + r : -1;
      cnt[(size_t)y * w + x] = 0;
    }
    __syncthreads();
  }
}

// Column pass: unite the runs of vertically connected pixels; a pair is skipped when the pair to its left already
// joined the same two runs.
__global__ void sgbm_cc_merge_kernel(const int16_t* __restrict__ d, int w, int h, int new_val, int max_diff,
                                     int* __restrict__ lab) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w || y + 1 >= h) return;
  const int i = y * w + x;
  const int v = d[i], u = d[i + w];
  if (v == new_val || u == new_val || abs(v - u) > max_diff) return;
  if (x > 0) {
    const int vl = d[i - 1], ul = d[i + w - 1];
    if (vl != new_val && ul != new_val && abs(v - vl) <= max_diff && abs(u - ul) <= max_diff && abs(vl - ul) <= max_diff)
      return;
  }
  sg_union(lab, lab[i], lab[i + w]);
}

// Every valid pixel looks up its root; one atomicAdd per distinct root per warp.
__global__ void sgbm_cc_count_kernel(int n, int* __restrict__ lab, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = i < n ? lab[i] : -1;
  int r = -1;
  if (l >= 0) r = sg_find(lab, l);
  const unsigned m = __match_any_sync(0xffffffffu, r);
  if (r >= 0) {
    if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(cnt + r, __popc(m));
    lab[i] = r;     // roots are the smallest index of their tree: overwriting a label with its root keeps every chain valid
  }
}

__global__ void sgbm_cc_apply_kernel(int n, const int* __restrict__ lab, const int* __restrict__ cnt, int max_size,
                                     int new_val, int16_t* __restrict__ d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = lab[i];
  if (r < 0) return;
  // a concurrent path-halving write of the count pass may have left an ancestor here instead of the root
  for (int q = lab[r]; q != r; q = lab[r]) r = q;
  if (cnt[r] <= max_size) d[i] = (int16_t)new_val;
}

// ------------------------------------------------------------------------------------ K_reproj
__global__ void sgbm_reproject_kernel(const int16_t* __restrict__ disp, int w, int h, const double* __restrict__ Q,
                                      float3* __restrict__ xyz, uint8_t* __restrict__ keep) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const size_t i = (size_t)y * w + x;
  const double d = (double)(float)disp[i];
  const double fx = (double)x, fy = (double)y;
  double hm[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    // Matx product: s = 0; s += a(i,k) * b(k) for k = 0..3, no contraction
    double s = __dmul_rn(Q[4 * r], fx);
    s = __dadd_rn(0.0, s);
    s = __dadd_rn(s, __dmul_rn(Q[4 * r + 1], fy));
    s = __dadd_rn(s, __dmul_rn(Q[4 * r + 2], d));
    s = __dadd_rn(s, __dmul_rn(Q[4 * r + 3], 1.0));
    hm[r] = s;
  }
  // the Vec3f destination is assigned first (rounding to float), then divided by the double W
  const float X = (float)__ddiv_rn((double)(float)hm[0], hm[3]);
  const float Y = (float)__ddiv_rn((double)(float)hm[1], hm[3]);
  const float Z = (float)__ddiv_rn((double)(float)hm[2], hm[3]);
  // StereoCV.cpp:240: skipped when z > 5 or z <= 0.01 (a NaN z is therefore kept, as in the reference)
  const bool skip = (Z > 5.0f) || (Z <= 0.01f);
  keep[i] = skip ? 0 : 1;
  xyz[i] = make_float3(X, __fmul_rn(Y, -1.0f), Z);
}

}  // namespace vo

// ====================================================================================== C ABI
using namespace vo;

#define SG_CHECK_CTX(c)                   \
  if (!(c)) return VO_ERR_INVALID_ARG;    \
  VO_CUDA(cudaSetDevice((c)->device))

void vo_sgbm_default_params(vo_sgbm_params* p) {
  if (!p) return;
  // StereoSGBM::create(1, 96, 7, 8*3, 32*3, 0, 60, 0, 3000, 5), reference src/StereoCV.cpp:39-50
  p->min_disparity = 1;
  p->num_disparities = 96;
  p->block_size = 7;
  p->p1 = 24;
  p->p2 = 96;
  p->disp12_max_diff = 0;
  p->pre_filter_cap = 60;
  p->uniqueness_ratio = 0;
  p->speckle_window_size = 3000;
  p->speckle_range = 5;
}

struct SgResolved {
  int minD, maxD, D, R, P1, P2, uniq, d12, ftzero, minX1, maxX1, W1, inv;
};

static int sgbm_resolve(const vo_sgbm_params* p, int w, int h, SgResolved& r) {
  if (!p || w < 3 || h < 1 || w > 16384 || h > 16384) return VO_ERR_INVALID_ARG;
  r.minD = p->min_disparity;
  r.D = p->num_disparities;
  if (r.D <= 0 || r.D % 16 != 0 || r.D > SG_MAX_D) {
    set_error("numDisparities %d: must be a positive multiple of 16, at most %d", r.D, SG_MAX_D);
    return VO_ERR_INVALID_ARG;
  }
  r.maxD = r.minD + r.D;
  const int block = p->block_size > 0 ? p->block_size : 5;
  if (block % 2 == 0 || block > 2 * SG_MAX_R + 1) {
    set_error("blockSize %d: must be odd and at most %d", block, 2 * SG_MAX_R + 1);
    return VO_ERR_INVALID_ARG;
  }
  r.R = block / 2;
  r.P1 = p->p1 > 0 ? p->p1 : 2;
  r.P2 = std::max(p->p2 > 0 ? p->p2 : 5, r.P1 + 1);
  r.uniq = p->uniqueness_ratio >= 0 ? p->uniqueness_ratio : 10;
  r.d12 = p->disp12_max_diff > 0 ? p->disp12_max_diff : 1;
  if (p->pre_filter_cap < 0 || p->pre_filter_cap > 126) return VO_ERR_INVALID_ARG;
  r.ftzero = std::max(p->pre_filter_cap, 15) | 1;
  // the int16 cost range OpenCV relies on: block^2 * (largest pixel cost) + P2 must fit
  if ((long long)block * block * (2 * r.ftzero + 63) + r.P2 > 32767 || r.uniq > 100) {
    set_error("blockSize / preFilterCap / P2 leave the int16 cost range");
    return VO_ERR_INVALID_ARG;
  }
  if (!(w - r.maxD > r.R)) {
    // cv2 throws here: "input images are too small for your window size and max disparity" (stereosgbm.cpp:511)
    set_error("width - (minDisparity + numDisparities) = %d must exceed blockSize / 2 = %d", w - r.maxD, r.R);
    return VO_ERR_INVALID_ARG;
  }
  r.minX1 = std::max(r.maxD, 0);
  r.maxX1 = w + std::min(r.minD, 0);
  r.W1 = r.maxX1 - r.minX1;
  r.inv = (r.minD - 1) * 16;
  return VO_OK;
}

static int sgbm_ensure(vo_ctx* c, int w, int h, const SgResolved& r, bool need_bgr) {
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  if (!s) {
    s = new Sgbm();
    c->sgbm = s;
    for (auto& e : s->ev) VO_CUDA(cudaEventCreate(&e));
    VO_CUDA(cudaEventCreateWithFlags(&s->ev_h, cudaEventDisableTiming));
    VO_CUDA(cudaEventCreateWithFlags(&s->ev_d, cudaEventDisableTiming));
    VO_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    VO_CUDA(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
    VO_CUDA(cudaStreamCreateWithFlags(&s->stream3, cudaStreamNonBlocking));
    VO_CUDA(cudaMalloc(&s->d_n, 4 * sizeof(int)));
    VO_CUDA(cudaMalloc(&s->dQ, 16 * sizeof(double)));
  }
  const size_t npx = (size_t)w * h;
  const size_t cost = (size_t)std::max(r.W1, 0) * h * r.D;
  if (npx > (size_t)s->w * s->h || cost > s->cost_elems) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    sgbm_release_buffers(s);
    if (s->gexec) {
      cudaGraphExecDestroy(s->gexec);
      s->gexec = nullptr;
    }
    s->have_disp = false;
    s->w = w;
    s->h = h;
    s->cost_elems = cost;
    for (int k = 0; k < 2; k++) VO_CUDA(cudaMalloc(&s->img[k], npx));
    VO_CUDA(cudaMalloc(&s->pl, 4 * npx * sizeof(uchar4)));
    VO_CUDA(cudaMalloc(&s->hsum, cost * 2 + 64));
    VO_CUDA(cudaMalloc(&s->C, cost * 2 + 64));
    VO_CUDA(cudaMalloc(&s->S, cost * 2 + 64));
    for (int k = 0; k < 3; k++) VO_CUDA(cudaMalloc(&s->vol[k], cost * 2 + 64));
    VO_CUDA(cudaMalloc(&s->key2, npx * sizeof(unsigned)));
    VO_CUDA(cudaMalloc(&s->rec, npx * sizeof(uint2)));
    for (int k = 0; k < 3; k++) VO_CUDA(cudaMalloc(&s->disp[k], npx * sizeof(int16_t)));
    VO_CUDA(cudaMalloc(&s->label, npx * sizeof(int)));
    VO_CUDA(cudaMalloc(&s->count, npx * sizeof(int)));
    VO_CUDA(cudaMalloc(&s->xyz, npx * sizeof(float3)));
    VO_CUDA(cudaMalloc(&s->keep, npx));
    VO_CUDA(cudaMalloc(&s->xyz_out, npx * sizeof(float3)));
    VO_CUDA(cudaMalloc(&s->idx_out, npx * sizeof(int)));
    size_t t1 = 0, t2 = 0;
    VO_CUDA(cub::DeviceSelect::Flagged(nullptr, t1, (const float3*)nullptr, (const uint8_t*)nullptr, (float3*)nullptr,
                                       (int*)nullptr, (int)npx, c->stream));
    VO_CUDA(cub::DeviceSelect::Flagged(nullptr, t2, cub::CountingInputIterator<int>(0), (const uint8_t*)nullptr,
                                       (int*)nullptr, (int*)nullptr, (int)npx, c->stream));
    s->cub_bytes = std::max(t1, t2);
    VO_CUDA(cudaMalloc(&s->cub_tmp, s->cub_bytes + 256));
  }
  if (need_bgr && !s->bgr[0]) {
    for (int k = 0; k < 2; k++) VO_CUDA(cudaMalloc(&s->bgr[k], 3 * (size_t)s->w * s->h));
  }
  return VO_OK;
}

template <int DPL, bool NARROW>
static int sgbm_path_launch(vo_ctx* c, cudaStream_t st, const int16_t* C, void* out_a, void* out_b, const SgResolved& r,
                            int h, int dir_a, int dir_b, int ndirs, int npaths) {
  const size_t smem = (size_t)SG_WPB * SgVec<DPL>::STAGES * 32 * 2 * DPL;
  c->launch_count++;
  sgbm_path_kernel<DPL, NARROW><<<dim3(div_up(npaths, SG_WPB), ndirs), SG_WPB * 32, smem, st>>>(
      C, out_a, out_b, r.W1, h, r.D, r.P1, r.P2, dir_a, dir_b, npaths);
  return VO_OK;
}

template <bool NARROW>
static int sgbm_path2_launch(vo_ctx* c, cudaStream_t st, const int16_t* C, void* out_a, void* out_b, const SgResolved& r, int h,
                             int dir_a, int dir_b, int ndirs, int npaths) {
  const size_t smem = (size_t)SG_WPB * 8 * 32 * 16;
  c->launch_count++;
  sgbm_path2_kernel<NARROW><<<dim3(div_up(npaths, 2 * SG_WPB), ndirs), SG_WPB * 32, smem, st>>>(
      C, out_a, out_b, r.W1, h, r.D, r.P1, r.P2, dir_a, dir_b, npaths);
  return VO_OK;
}

// All five directions run at once, each into its own volume: the horizontal pair (the long paths: W1 steps, only
// 2 * h warps) on a second stream, the diagonal pair on a third, the vertical direction on the main stream; none of
// them alone fills the machine (a path is one warp), together they stream C five times at HBM speed.  The
// winner-take-all kernel then sums the five volumes per pixel.
template <int DPL, bool NARROW>
static int sgbm_paths(vo_ctx* c, Sgbm* s, const SgResolved& r, int h, const SgWta& wta) {
  void* L0 = s->hsum;   // hsum is dead once C exists
  const int nd = r.W1 + h - 1;
  VO_CUDA(cudaEventRecord(s->ev_fork, c->stream));
  VO_CUDA(cudaStreamWaitEvent(s->stream2, s->ev_fork, 0));
  VO_CUDA(cudaStreamWaitEvent(s->stream3, s->ev_fork, 0));
  VO_TRY((sgbm_path_launch<DPL, NARROW>(c, s->stream2, s->C, L0, s->vol[0], r, h, 0, 4, 2, h)));
  VO_CUDA(cudaEventRecord(s->ev_h, s->stream2));
  static const bool one_path = getenv("VO_B200_SGBM_ONE_PATH_PER_WARP") != nullptr;   // experiment switch
  const bool two = r.D <= 128 && !one_path;
  if (two) VO_TRY((sgbm_path2_launch<NARROW>(c, s->stream3, s->C, s->vol[1], s->vol[2], r, h, 1, 3, 2, nd)));
  else VO_TRY((sgbm_path_launch<DPL, NARROW>(c, s->stream3, s->C, s->vol[1], s->vol[2], r, h, 1, 3, 2, nd)));
  VO_CUDA(cudaEventRecord(s->ev_d, s->stream3));
  if (two) VO_TRY((sgbm_path2_launch<NARROW>(c, c->stream, s->C, s->S, nullptr, r, h, 2, 2, 1, r.W1)));
  else VO_TRY((sgbm_path_launch<DPL, NARROW>(c, c->stream, s->C, s->S, nullptr, r, h, 2, 2, 1, r.W1)));
  if (s->stage_events) VO_CUDA(cudaEventRecord(s->ev[4], c->stream));
  VO_CUDA(cudaStreamWaitEvent(c->stream, s->ev_h, 0));
  VO_CUDA(cudaStreamWaitEvent(c->stream, s->ev_d, 0));
  if (s->stage_events) VO_CUDA(cudaEventRecord(s->ev[5], c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    sgbm_wta_kernel<DPL, NARROW><<<dim3(div_up(r.W1, 8 * SG_WTA_PIX), h), 256, 0, c->stream>>>(s->C, L0, s->vol[0], s->vol[1],
                                                                                             s->vol[2], s->S, r.W1, h, r.D, wta);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// device pipeline on s->img[0..1] (tight gray) -> s->disp[2]
// enqueues the whole pipeline on the context's streams (also under stream capture, see sgbm_run)
static int sgbm_enqueue(vo_ctx* c, int w, int h, const vo_sgbm_params* p, const SgResolved& r, bool stage_events) {
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  const size_t npx = (size_t)w * h;
  const int n = (int)npx;
  s->stage_events = stage_events;
  {
    LaunchScope ls(c, VO_K_MISC);
    sgbm_fill_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(s->rec, s->key2, npx);
  }
  if (r.W1 > 0) {
    {
      LaunchScope ls(c, VO_K_MISC);
      sgbm_prefilter_kernel<<<dim3(div_up(w, 128), h, 2), 128, 0, c->stream>>>(s->img[0], s->img[1], w, h, r.ftzero, s->pl);
    }
    if (s->stage_events) VO_CUDA(cudaEventRecord(s->ev[2], c->stream));
    {
      LaunchScope ls(c, VO_K_MISC);
      const size_t smem = (size_t)(SG_TX + 2 * r.R) * r.D * sizeof(uint16_t) + 2 * (size_t)(SG_TX + 2 * r.R + r.D - 1) * sizeof(uchar4);
      sgbm_hsum_kernel<<<dim3(div_up(r.W1, SG_TX), h), 256, smem, c->stream>>>(s->pl, w, h, r.W1, r.D, r.minD, r.minX1, r.R,
                                                                             s->hsum);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      const size_t row_vecs = (size_t)r.W1 * r.D / 8;     // D is a multiple of 16
      const int strips = 16, rps = div_up(h, strips);
      sgbm_vsum_kernel<<<dim3((unsigned)((row_vecs + 255) / 256), strips), 256, 0, c->stream>>>(
          reinterpret_cast<const uint4*>(s->hsum), h, row_vecs, r.R, rps, reinterpret_cast<uint4*>(s->C));
    }
    if (s->stage_events) VO_CUDA(cudaEventRecord(s->ev[3], c->stream));
    SgWta wta{s->rec, s->key2, w, r.minD, r.minX1, r.uniq};
    static const bool wide = getenv("VO_B200_SGBM_WIDE") != nullptr;   // force the int16 volumes (test switch)
    const bool narrow = r.P2 <= 255 && !wide;
    if (r.D <= 128) {
      if (narrow) VO_TRY((sgbm_paths<4, true>(c, s, r, h, wta)));
      else VO_TRY((sgbm_paths<4, false>(c, s, r, h, wta)));
    } else {
      if (narrow) VO_TRY((sgbm_paths<8, true>(c, s, r, h, wta)));
      else VO_TRY((sgbm_paths<8, false>(c, s, r, h, wta)));
    }
  } else {
    for (int k = 2; k <= 5; k++)
      if (s->stage_events) VO_CUDA(cudaEventRecord(s->ev[k], c->stream));
  }
  if (s->stage_events) VO_CUDA(cudaEventRecord(s->ev[6], c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    static const bool keep_raw = getenv("VO_B200_SGBM_DEBUG") != nullptr;   // stage 1 of vo_debug_sgbm_stage
    sgbm_lrcheck_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(s->rec, s->key2, w, h, r.D, r.minD, r.minX1, r.maxX1,
                                                                       r.d12, s->disp[1], keep_raw ? s->disp[0] : nullptr);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    sgbm_median3_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(s->disp[1], w, h, s->disp[2]);
  }
  if (s->stage_events) VO_CUDA(cudaEventRecord(s->ev[7], c->stream));
  if (p->speckle_window_size > 0) {
    const int max_diff = 16 * p->speckle_range;
    {
      LaunchScope ls(c, VO_K_MISC);
      sgbm_cc_rows_kernel<<<h, 256, 0, c->stream>>>(s->disp[2], w, r.inv, max_diff, s->label, s->count);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      sgbm_cc_merge_kernel<<<dim3(div_up(w, 128), h), 128, 0, c->stream>>>(s->disp[2], w, h, r.inv, max_diff, s->label);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      sgbm_cc_count_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(n, s->label, s->count);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      sgbm_cc_apply_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(n, s->label, s->count, p->speckle_window_size, r.inv,
                                                                 s->disp[2]);
    }
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// The pipeline is ~17 short launches on three streams: for a repeated (size, parameters) it is captured once into a
// CUDA graph and replayed with one launch, which removes the launch gaps between the small kernels and most of the
// host time of a call.  The graph dies with the buffers it points to (sgbm_ensure).
static int sgbm_run(vo_ctx* c, int w, int h, const vo_sgbm_params* p, const SgResolved& r) {
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  static const bool no_graph = getenv("VO_B200_SGBM_NO_GRAPH") != nullptr;
  s->have_disp = true;
  s->last_w = w;
  s->last_h = h;
  VO_CUDA(cudaEventRecord(s->ev[1], c->stream));
  // per-stage events only exist on plain launches: profiling (vo_profile_enable) selects them
  if (no_graph || s->graph_failed || c->prof.mask) {
    VO_TRY(sgbm_enqueue(c, w, h, p, r, true));
    s->staged = true;
    VO_CUDA(cudaEventRecord(s->ev[8], c->stream));
    return VO_OK;
  }
  const bool same = s->gexec && s->g_w == w && s->g_h == h && memcmp(&s->g_params, p, sizeof(*p)) == 0;
  if (!same) {
    if (s->gexec) {
      cudaGraphExecDestroy(s->gexec);
      s->gexec = nullptr;
    }
    const int64_t before = c->launch_count;
    VO_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = sgbm_enqueue(c, w, h, p, r, false);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    s->g_launches = (int)(c->launch_count - before);
    c->launch_count = before;
    if (rc != VO_OK || e != cudaSuccess || !g || cudaGraphInstantiate(&s->gexec, g, 0) != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      s->gexec = nullptr;
      s->graph_failed = true;            // this driver / configuration cannot capture it: plain launches from now on
      VO_CUDA(cudaEventRecord(s->ev[1], c->stream));
      VO_TRY(sgbm_enqueue(c, w, h, p, r, true));
      s->staged = true;
      VO_CUDA(cudaEventRecord(s->ev[8], c->stream));
      return VO_OK;
    }
    cudaGraphDestroy(g);
    s->g_w = w;
    s->g_h = h;
    s->g_params = *p;
  }
  VO_CUDA(cudaGraphLaunch(s->gexec, c->stream));
  c->launch_count += s->g_launches;
  s->staged = false;
  VO_CUDA(cudaEventRecord(s->ev[8], c->stream));
  return VO_OK;
}

// Caller buffers are usually pageable (cv::Mat / std::vector memory): a pageable cudaMemcpy runs at a few GB/s, so
// such buffers go through pinned staging owned by the context (one host memcpy + one DMA); pinned, managed and
// device pointers are copied directly.
static bool sg_needs_staging(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

static int sg_stage_reserve(uint8_t** buf, size_t* have, size_t need) {
  if (need <= *have) return VO_OK;
  cudaFreeHost(*buf);
  *buf = nullptr;
  *have = 0;
  VO_CUDA(cudaMallocHost(buf, need));
  *have = need;
  return VO_OK;
}

// host image (row_bytes per row, stride between rows) -> tight device buffer
static int sg_upload(vo_ctx* c, Sgbm* s, int k, uint8_t* d_dst, const uint8_t* src, int stride, size_t row_bytes, int h) {
  if (!sg_needs_staging(src)) {
    VO_CUDA(cudaMemcpy2DAsync(d_dst, row_bytes, src, stride, row_bytes, h, cudaMemcpyDefault, c->stream));
    return VO_OK;
  }
  const size_t bytes = row_bytes * h;
  if (bytes > s->h_in_bytes) {
    VO_CUDA(cudaStreamSynchronize(c->stream));
    size_t have0 = s->h_in_bytes, have1 = s->h_in_bytes;
    VO_TRY(sg_stage_reserve(&s->h_in[0], &have0, bytes));
    VO_TRY(sg_stage_reserve(&s->h_in[1], &have1, bytes));
    s->h_in_bytes = bytes;
  }
  if ((size_t)stride == row_bytes) {
    memcpy(s->h_in[k], src, bytes);
  } else {
    for (int y = 0; y < h; y++) memcpy(s->h_in[k] + (size_t)y * row_bytes, src + (size_t)y * stride, row_bytes);
  }
  VO_CUDA(cudaMemcpyAsync(d_dst, s->h_in[k], bytes, cudaMemcpyHostToDevice, c->stream));
  return VO_OK;
}

static int sgbm_finish(vo_ctx* c, int w, int h, int16_t* disp, int disp_stride) {
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  const size_t row = (size_t)w * 2;
  bool staged = false;
  if (disp) {
    if (sg_needs_staging(disp)) {
      VO_TRY(sg_stage_reserve(&s->h_out, &s->h_out_bytes, row * h));
      VO_CUDA(cudaMemcpyAsync(s->h_out, s->disp[2], row * h, cudaMemcpyDeviceToHost, c->stream));
      staged = true;
    } else {
      VO_CUDA(cudaMemcpy2DAsync(disp, disp_stride, s->disp[2], row, row, h, cudaMemcpyDefault, c->stream));
    }
  }
  VO_CUDA(cudaEventRecord(s->ev[9], c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  if (staged) {
    if ((size_t)disp_stride == row) memcpy(disp, s->h_out, row * h);
    else
      for (int y = 0; y < h; y++) memcpy((uint8_t*)disp + (size_t)y * disp_stride, s->h_out + (size_t)y * row, row);
  }
  auto span = [&](int a, int b) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s->ev[a], s->ev[b]) != cudaSuccess) {
      cudaGetLastError();
      ms = 0.f;
    }
    return ms;
  };
  for (int k = 0; k < 9; k++) s->ms[k] = 0.f;
  s->ms[0] = span(0, 1);
  s->ms[8] = span(8, 9);
  s->pipeline_ms = span(1, 8);
  if (s->staged)
    for (int k = 1; k < 8; k++) s->ms[k] = span(k, k + 1);
  return VO_OK;
}

int vo_sgbm_compute(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride, int width, int height,
                    const vo_sgbm_params* p, int16_t* disp, int disp_stride) {
  SG_CHECK_CTX(c);
  if (!left || !right || !p || stride < width || (disp && disp_stride < 2 * width)) return VO_ERR_INVALID_ARG;
  SgResolved r;
  VO_TRY(sgbm_resolve(p, width, height, r));
  VO_TRY(sgbm_ensure(c, width, height, r, false));
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  VO_CUDA(cudaEventRecord(s->ev[0], c->stream));
  VO_TRY(sg_upload(c, s, 0, s->img[0], left, stride, (size_t)width, height));
  VO_TRY(sg_upload(c, s, 1, s->img[1], right, stride, (size_t)width, height));
  VO_TRY(sgbm_run(c, width, height, p, r));
  return sgbm_finish(c, width, height, disp, disp_stride);
}

int vo_stereo_match(vo_ctx* c, const uint8_t* left_bgr, const uint8_t* right_bgr, int stride, int width, int height,
                    const vo_sgbm_params* p, int16_t* disp, int disp_stride) {
  SG_CHECK_CTX(c);
  if (!left_bgr || !right_bgr || !p || stride < 3 * width || (disp && disp_stride < 2 * width)) return VO_ERR_INVALID_ARG;
  SgResolved r;
  VO_TRY(sgbm_resolve(p, width, height, r));
  VO_TRY(sgbm_ensure(c, width, height, r, true));
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  VO_CUDA(cudaEventRecord(s->ev[0], c->stream));
  const uint8_t* src[2] = {left_bgr, right_bgr};
  for (int k = 0; k < 2; k++) {
    VO_TRY(sg_upload(c, s, k, s->bgr[k], src[k], stride, 3 * (size_t)width, height));
    // cvtColor(BGR2GRAY), src/StereoCV.cpp:35-36
    VO_TRY(bgr2gray_launch_wh(c, s->bgr[k], 3 * width, s->img[k], width, width, height));
  }
  VO_TRY(sgbm_run(c, width, height, p, r));
  return sgbm_finish(c, width, height, disp, disp_stride);
}

int vo_sgbm_timing(vo_ctx* c, float ms[9], float* pipeline_ms) {
  if (!c || !ms || !c->sgbm) return VO_ERR_INVALID_ARG;
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  for (int k = 0; k < 9; k++) ms[k] = s->ms[k];
  if (pipeline_ms) *pipeline_ms = s->pipeline_ms;
  return VO_OK;
}

int vo_debug_sgbm_stage(vo_ctx* c, int stage, void* out, uint64_t bytes) {
  SG_CHECK_CTX(c);
  if (!out || !c->sgbm) return VO_ERR_INVALID_ARG;
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  if (!s->have_disp) return VO_ERR_INVALID_ARG;
  const void* src = nullptr;
  size_t avail = 0;
  const size_t npx = (size_t)s->last_w * s->last_h;
  switch (stage) {
    case 0: src = s->C; avail = s->cost_elems * 2; break;
    case 1: src = s->disp[0]; avail = npx * 2; break;     // winner-take-all + sub-pixel, before the left-right check
                                                          // (kept only when VO_B200_SGBM_DEBUG is set)
    case 2: src = s->disp[1]; avail = npx * 2; break;     // after the left-right check
    case 3: src = s->pl; avail = 4 * npx * 4; break;      // prefilter planes
    default: return VO_ERR_INVALID_ARG;
  }
  if (bytes > avail) return VO_ERR_CAPACITY;
  VO_CUDA(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

int vo_reproject_disparity(vo_ctx* c, const int16_t* disp, int disp_stride, int width, int height, const double Q[16],
                           float* xyz, int32_t* pix_idx, int cap, int* n_out) {
  SG_CHECK_CTX(c);
  if (!Q || !n_out || cap < 0 || width < 1 || height < 1 || (cap > 0 && !xyz)) return VO_ERR_INVALID_ARG;
  *n_out = 0;
  Sgbm* s = reinterpret_cast<Sgbm*>(c->sgbm);
  if (disp) {
    if (disp_stride < 2 * width) return VO_ERR_INVALID_ARG;
    SgResolved r{};
    r.W1 = 0;
    r.D = 16;
    VO_TRY(sgbm_ensure(c, width, height, r, false));
    s = reinterpret_cast<Sgbm*>(c->sgbm);
    VO_TRY(sg_upload(c, s, 0, reinterpret_cast<uint8_t*>(s->disp[2]), reinterpret_cast<const uint8_t*>(disp), disp_stride,
                     (size_t)width * 2, height));
    s->have_disp = true;
    s->last_w = width;
    s->last_h = height;
  } else if (!s || !s->have_disp || s->last_w != width || s->last_h != height) {
    set_error("vo_reproject_disparity: no device-resident disparity of this size (run vo_sgbm_compute first)");
    return VO_ERR_INVALID_ARG;
  }
  const int n = width * height;
  double* dQ = s->dQ;
  VO_CUDA(cudaMemcpyAsync(dQ, Q, 16 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    sgbm_reproject_kernel<<<dim3(div_up(width, 128), height), 128, 0, c->stream>>>(s->disp[2], width, height, dQ, s->xyz,
                                                                                  s->keep);
  }
  size_t tb = s->cub_bytes;
  VO_CUDA(cub::DeviceSelect::Flagged(s->cub_tmp, tb, s->xyz, s->keep, s->xyz_out, s->d_n, n, c->stream));
  tb = s->cub_bytes;
  VO_CUDA(cub::DeviceSelect::Flagged(s->cub_tmp, tb, cub::CountingInputIterator<int>(0), s->keep, s->idx_out, s->d_n + 1, n,
                                     c->stream));
  c->launch_count += 2;
  int hn[2] = {0, 0};
  VO_CUDA(cudaMemcpyAsync(hn, s->d_n, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  const int m = hn[0];
  *n_out = m;
  const int k = std::min(m, cap);
  if (k > 0) {
    const size_t bx = (size_t)k * sizeof(float3), bi = pix_idx ? (size_t)k * sizeof(int) : 0;
    if (sg_needs_staging(xyz) || (pix_idx && sg_needs_staging(pix_idx))) {
      // pageable destination: DMA into pinned staging in chunks, each chunk's host copy overlapping the next DMA
      VO_TRY(sg_stage_reserve(&s->h_out, &s->h_out_bytes, bx + bi));
      constexpr size_t CH = 1 << 20;
      constexpr int NEV = 10;
      const size_t total = bx + bi;
      const int nch = (int)((total + CH - 1) / CH);
      for (int ch = 0; ch < nch; ch++) {
        const size_t o = (size_t)ch * CH, len = std::min(CH, total - o);
        // staging layout: points first, indices behind them; a chunk may straddle the two device arrays
        if (o < bx) {
          const size_t l1 = std::min(len, bx - o);
          VO_CUDA(cudaMemcpyAsync(s->h_out + o, reinterpret_cast<const uint8_t*>(s->xyz_out) + o, l1, cudaMemcpyDeviceToHost,
                                  c->stream));
          if (l1 < len)
            VO_CUDA(cudaMemcpyAsync(s->h_out + bx, s->idx_out, len - l1, cudaMemcpyDeviceToHost, c->stream));
        } else {
          VO_CUDA(cudaMemcpyAsync(s->h_out + o, reinterpret_cast<const uint8_t*>(s->idx_out) + (o - bx), len,
                                  cudaMemcpyDeviceToHost, c->stream));
        }
        if (ch < NEV) VO_CUDA(cudaEventRecord(s->ev[ch], c->stream));
      }
      for (int ch = 0; ch < nch; ch++) {
        if (ch < NEV) VO_CUDA(cudaEventSynchronize(s->ev[ch]));
        else VO_CUDA(cudaStreamSynchronize(c->stream));
        const size_t o = (size_t)ch * CH, len = std::min(CH, total - o);
        if (o < bx) {
          const size_t l1 = std::min(len, bx - o);
          memcpy(reinterpret_cast<uint8_t*>(xyz) + o, s->h_out + o, l1);
          if (l1 < len) memcpy(pix_idx, s->h_out + bx, len - l1);
        } else {
          memcpy(reinterpret_cast<uint8_t*>(pix_idx) + (o - bx), s->h_out + o, len);
        }
      }
    } else {
      VO_CUDA(cudaMemcpyAsync(xyz, s->xyz_out, bx, cudaMemcpyDefault, c->stream));
      if (bi) VO_CUDA(cudaMemcpyAsync(pix_idx, s->idx_out, bi, cudaMemcpyDefault, c->stream));
      VO_CUDA(cudaStreamSynchronize(c->stream));
    }
  }
  return m > cap ? VO_ERR_CAPACITY : VO_OK;
}
