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
// The second pass (mean / standard deviation of those means in double, in input order, and the threshold
// test) is sequential by definition and runs on the host over N floats.
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

int sor_mean_knn_launch(vo_ctx* c, const float3* d_pts, int n, int mean_k, float* d_mean) {
  if (n <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_MISC);
    sor_mean_knn_kernel<<<div_up(n, SOR_WARPS), SOR_WARPS * 32, 0, c->stream>>>(d_pts, n, mean_k + 1, d_mean);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
