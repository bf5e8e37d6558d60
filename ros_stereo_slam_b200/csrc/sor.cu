// sor.cu -- statistical outlier removal of the keyframe / per-frame point cloud:
// visualSLAM::SORcloud (reference src/rosFuncs.cpp:9-39; called every frame at src/VisualSLAM.cpp:154 and
// on keyframes at :128) = drop points with -z > 500, then pcl::StatisticalOutlierRemoval with
// meanK = 200, stddevMulThresh = 0.01.  SURVEY.md section 8(f)-3.
//
// The part that costs time is the first pass of PCL's filter: for every point the mean distance to its
// meanK nearest neighbours (exact kNN; PCL asks FLANN for meanK+1 neighbours and skips the query itself).
// Here: one warp per query point, brute force over all points (N <= a few 10^4: 3 subtractions and 3
// multiply-adds per pair, no tree), and an exact radix select over the float bit patterns of the squared
// distances (non-negative floats order like unsigned integers): four 8-bit passes find the (meanK+1)-th
// smallest squared distance T and how many of the ties at T belong to the neighbour set, a fifth pass sums
// sqrtf(d2) over d2 < T.  The candidates stream through a shared-memory tile shared by the warps of a CTA.
// Squared distances are accumulated like FLANN's L2_Simple<float> (((dx*dx) + dy*dy) + dz*dz in float, no
// FMA contraction); the square roots are float, their sum double, the mean is cast to float -- PCL's
// arithmetic.  The query itself contributes the one 0 that PCL skips (k = 0).
// Pruning (sor_mean_knn_sorted_kernel, the default): the points are first sorted by z (CUB radix sort).  Any
// meanK+1 candidates bound T from above, so the (meanK+1)-th smallest distance among the 512 neighbours of the
// query in z order gives U >= T, and every true neighbour lies in the slab |z - z_q| <= sqrt(U), a contiguous
// index range found by binary search.  The same exact radix select then runs over the slab only (a few per cent
// of the cloud) -- same candidate semantics, same arithmetic, bit-identical result, ~4x less time.  The
// unpruned kernel stays as the reference implementation (VO_B200_SOR_UNPRUNED=1) and for tiny clouds.
// The second pass (mean / standard deviation of those means in double, in input order, and the threshold
// test) is sequential by definition and runs on the host over N floats.
#include <cub/cub.cuh>

#include "common.cuh"

namespace vo {

constexpr int SOR_WARPS = 8;
constexpr int SOR_TILE = 1024;   // candidates per shared-memory tile (12 KB)

__device__ __forceinline__ unsigned sor_key(float qx, float qy, float qz, float3 c) {
  const float dx = __fsub_rn(qx, c.x), dy = __fsub_rn(qy, c.y), dz = __fsub_rn(qz, c.z);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __float_as_uint(d2);
}

__global__ void __launch_bounds__(SOR_WARPS * 32)
sor_mean_knn_kernel(const float3* __restrict__ pts, int n, int k_plus_1, float* __restrict__ mean_dist) {
  __shared__ float3 tile[SOR_TILE];
  __shared__ int hist[SOR_WARPS][256];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * SOR_WARPS + wid;
  const bool active = q < n;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const float3 p = pts[q];
    qx = p.x; qy = p.y; qz = p.z;
  }
  unsigned prefix = 0;        // bits of T found so far
  int rank = k_plus_1;        // rank of T among the keys that share `prefix`
  // ---- four radix passes, most significant byte first
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    for (int i = lane; i < 256; i += 32) hist[wid][i] = 0;
    __syncwarp();
    for (int t0 = 0; t0 < n; t0 += SOR_TILE) {
      __syncthreads();
      for (int i = threadIdx.x; i < SOR_TILE && t0 + i < n; i += blockDim.x) tile[i] = pts[t0 + i];
      __syncthreads();
      const int cnt = min(SOR_TILE, n - t0);
      if (active) {
        for (int i = lane; i < cnt; i += 32) {
          const unsigned key = sor_key(qx, qy, qz, tile[i]);
          // pass 0 has no prefix; a shift by 32 is undefined, so handle it explicitly
          const bool match = pass == 0 || (key >> (shift + 8)) == prefix;
          if (match) atomicAdd(&hist[wid][(key >> shift) & 255u], 1);
        }
      }
    }
    __syncwarp();
    if (active) {
      // bucket b with cumulative count >= rank: 8 buckets per lane, warp prefix sum
      int c[8], s = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        c[j] = hist[wid][lane * 8 + j];
        s += c[j];
      }
      int incl = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const int excl = incl - s;
      const bool mine = excl < rank && rank <= incl;     // exactly one lane (rank <= number of matching keys)
      int b = 0, before = 0;
      if (mine) {
        int run = excl;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (run < rank && rank <= run + c[j]) {
            b = lane * 8 + j;
            before = run;
          }
          run += c[j];
        }
      }
      const unsigned who = __ballot_sync(0xffffffffu, mine);
      const int src = __ffs(who) - 1;
      b = __shfl_sync(0xffffffffu, b, src);
      before = __shfl_sync(0xffffffffu, before, src);
      prefix = (prefix << 8) | (unsigned)b;
      rank -= before;
    }
    __syncwarp();
  }
  // prefix = key of the (k+1)-th smallest squared distance, rank = how many keys equal to it are taken
  // ---- fifth pass: sum of the square roots below T
  double sum = 0.0;
  for (int t0 = 0; t0 < n; t0 += SOR_TILE) {
    __syncthreads();
    for (int i = threadIdx.x; i < SOR_TILE && t0 + i < n; i += blockDim.x) tile[i] = pts[t0 + i];
    __syncthreads();
    const int cnt = min(SOR_TILE, n - t0);
    if (active) {
      for (int i = lane; i < cnt; i += 32) {
        const unsigned key = sor_key(qx, qy, qz, tile[i]);
        if (key < prefix) sum += (double)__fsqrt_rn(__uint_as_float(key));
      }
    }
  }
  if (active) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) {
      sum += (double)rank * (double)__fsqrt_rn(__uint_as_float(prefix));
      mean_dist[q] = (float)(sum / (double)(k_plus_1 - 1));
    }
  }
}

// ---------------------------------------------------------------------------------------------- pruned
constexpr int SORS_WARPS = 8;
constexpr int SORS_WINDOW = 512;

// one radix pass over candidates [lo, hi) of `pts`; keys above `cap_key` never count (they are beyond the bound)
__device__ __forceinline__ void sors_hist_pass(const float3* __restrict__ pts, int lo, int hi, float qx, float qy, float qz,
                                               int pass, unsigned prefix, unsigned cap_key, int* hist, int lane) {
  const int shift = 24 - 8 * pass;
  for (int i = lane; i < 256; i += 32) hist[i] = 0;
  __syncwarp();
  for (int i = lo + lane; i < hi; i += 32) {
    const unsigned key = sor_key(qx, qy, qz, pts[i]);
    if (key <= cap_key && (pass == 0 || (key >> (shift + 8)) == prefix)) atomicAdd(&hist[(key >> shift) & 255u], 1);
  }
  __syncwarp();
}

// bucket holding rank `rank` in hist (whole warp); updates prefix and rank
__device__ __forceinline__ void sors_pick(const int* hist, int lane, unsigned& prefix, int& rank) {
  int c[8], s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    c[j] = hist[lane * 8 + j];
    s += c[j];
  }
  int incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int excl = incl - s;
  const bool mine = excl < rank && rank <= incl;
  int b = 0, before = 0;
  if (mine) {
    int run = excl;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (run < rank && rank <= run + c[j]) {
        b = lane * 8 + j;
        before = run;
      }
      run += c[j];
    }
  }
  const unsigned who = __ballot_sync(0xffffffffu, mine);
  const int src = __ffs(who) - 1;
  b = __shfl_sync(0xffffffffu, b, src);
  before = __shfl_sync(0xffffffffu, before, src);
  prefix = (prefix << 8) | (unsigned)b;
  rank -= before;
  __syncwarp();
}

// pts: the cloud sorted by z; zs: the sorted z values; order[q]: original index of sorted point q
__global__ void __launch_bounds__(SORS_WARPS * 32)
sor_mean_knn_sorted_kernel(const float3* __restrict__ pts, const float* __restrict__ zs, const int* __restrict__ order, int n,
                           int k_plus_1, float* __restrict__ mean_dist) {
  __shared__ int hist_all[SORS_WARPS][256];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * SORS_WARPS + wid;
  if (q >= n) return;               // whole warps only; no block-level barrier below
  int* hist = hist_all[wid];
  const float3 p = pts[q];
  const float qx = p.x, qy = p.y, qz = p.z;
  // ---- upper bound U: (k+1)-th smallest among the z-order window around q
  const int m = min(n, max(SORS_WINDOW, 2 * k_plus_1));
  const int w_lo = max(0, min(q - m / 2, n - m));
  unsigned U = 0;
  int rank = k_plus_1;
  for (int pass = 0; pass < 4; pass++) {
    sors_hist_pass(pts, w_lo, w_lo + m, qx, qy, qz, pass, U, 0xffffffffu, hist, lane);
    sors_pick(hist, lane, U, rank);
  }
  // ---- slab: every point with squared distance <= U has |z - z_q| <= sqrt(U) (rounded up generously)
  const float rz = __fsqrt_ru(__uint_as_float(U)) * 1.00001f + 1e-30f;
  const float z_lo = qz - rz, z_hi = qz + rz;
  int lo = 0, hi = n;
  {
    int a = 0, b = n;                        // lower_bound(z_lo) in [0, n)
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (zs[mid] < z_lo) a = mid + 1; else b = mid;
    }
    lo = a;
    a = lo; b = n;                           // upper_bound(z_hi)
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (zs[mid] <= z_hi) a = mid + 1; else b = mid;
    }
    hi = a;
  }
  lo = min(lo, w_lo);                        // the window itself is always part of the candidate range
  hi = max(hi, w_lo + m);
  // ---- exact select over the slab (keys above U cannot be among the k+1 smallest)
  unsigned T = 0;
  rank = k_plus_1;
  for (int pass = 0; pass < 4; pass++) {
    sors_hist_pass(pts, lo, hi, qx, qy, qz, pass, T, U, hist, lane);
    sors_pick(hist, lane, T, rank);
  }
  double sum = 0.0;
  for (int i = lo + lane; i < hi; i += 32) {
    const unsigned key = sor_key(qx, qy, qz, pts[i]);
    if (key < T) sum += (double)__fsqrt_rn(__uint_as_float(key));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) {
    sum += (double)rank * (double)__fsqrt_rn(__uint_as_float(T));
    mean_dist[order[q]] = (float)(sum / (double)(k_plus_1 - 1));
  }
}

__global__ void sor_keys_kernel(const float3* __restrict__ pts, int n, float* __restrict__ z, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  z[i] = pts[i].z;
  idx[i] = i;
}

__global__ void sor_gather_kernel(const float3* __restrict__ pts, const int* __restrict__ order, int n, float3* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = pts[order[i]];
}

int sor_mean_knn_launch(vo_ctx* c, const float3* d_pts, int n, int mean_k, float* d_mean) {
  if (n <= 0) return VO_OK;
  static const bool unpruned = getenv("VO_B200_SOR_UNPRUNED") != nullptr;
  if (unpruned || n < 2 * SORS_WINDOW) {
    LaunchScope ls(c, VO_K_MISC);
    sor_mean_knn_kernel<<<div_up(n, SOR_WARPS), SOR_WARPS * 32, 0, c->stream>>>(d_pts, n, mean_k + 1, d_mean);
    VO_CUDA(cudaGetLastError());
    return VO_OK;
  }
  // scratch (allocated once, sized for max_points): z keys in/out, order in/out, sorted points, CUB temp
  const size_t cap = (size_t)c->cap;
  if (!c->d_sor) {
    size_t tmp = 0;
    VO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const float*)nullptr, (float*)nullptr, (const int*)nullptr,
                                            (int*)nullptr, (int)cap, 0, 32, c->stream));
    c->sor_tmp_bytes = tmp;
    VO_CUDA(cudaMalloc(&c->d_sor, cap * (4 * sizeof(float) + sizeof(float3)) + tmp + 256));
  }
  float* z_in = reinterpret_cast<float*>(c->d_sor);
  float* z_out = z_in + cap;
  int* ord_in = reinterpret_cast<int*>(z_out + cap);
  int* ord_out = ord_in + cap;
  float3* sorted = reinterpret_cast<float3*>(ord_out + cap);
  void* tmp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(sorted + cap) + 255) & ~uintptr_t(255));
  size_t tmp_bytes = c->sor_tmp_bytes;
  {
    LaunchScope ls(c, VO_K_MISC);
    sor_keys_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_pts, n, z_in, ord_in);
  }
  VO_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, z_in, z_out, ord_in, ord_out, n, 0, 32, c->stream));
  {
    LaunchScope ls(c, VO_K_MISC);
    sor_gather_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_pts, ord_out, n, sorted);
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    sor_mean_knn_sorted_kernel<<<div_up(n, SORS_WARPS), SORS_WARPS * 32, 0, c->stream>>>(sorted, z_out, ord_out, n, mean_k + 1,
                                                                                     d_mean);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
