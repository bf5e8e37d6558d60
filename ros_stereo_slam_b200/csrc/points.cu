// points.cu -- small per-point kernels of the hot path:
//   K9a grid keypoints      visualSLAM::denseKeypointExtractor   (reference src/tracking.cpp:4-12)
//   K3  ordered compaction  the push_back loops                  (src/tracking.cpp:20-27,35-42,66-72,78-84)
//   K5  DLT triangulation   cv::triangulatePoints + dehomogenise (src/triangulation.cpp:152-160)
//       + rigid transform   insertKeyFrames / update3dtransformation (src/keyFrameManagement.cpp:20-30,33-46)
#include "common.cuh"
#include "cvmath.cuh"

namespace vo {

// ---------------------------------------------------------------------------- grid
__global__ void grid_kernel(float2* __restrict__ xy, int nx, int ny, int step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx * ny) return;
  const int gy = i / nx, gx = i - gy * nx;
  xy[i] = make_float2((float)(step + gx * step), (float)(step + gy * step));
}

int grid_launch(vo_ctx* c, int rows, int cols, int step, float2* d_xy, int* n_out) {
  // for (y = s; y < rows - s; y += s) for (x = s; x < cols - s; x += s)
  int ny = 0, nx = 0;
  if (step > 0) {
    for (int y = step; y < rows - step; y += step) ny++;
    for (int x = step; x < cols - step; x += step) nx++;
  }
  const int n = nx * ny;
  *n_out = n;
  if (n == 0) return VO_OK;
  if (n > c->cap) {
    set_error("grid of %d keypoints exceeds max_points %d", n, c->cap);
    return VO_ERR_CAPACITY;
  }
  {
    LaunchScope ls(c, VO_K_MISC);
    grid_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_xy, nx, ny, step);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ---------------------------------------------------------------------------- compaction
// Single CTA, 1024 threads: each thread owns a contiguous chunk (order preserving), block
// exclusive scan of the chunk counts, then the ordered scatter.  N <= 131072 -> <= 128
// flags per thread; the arrays are L2-resident (<= 1.5 MB) so this is latency-, not
// bandwidth-bound.
constexpr int COMPACT_THREADS = 1024;

__global__ void __launch_bounds__(COMPACT_THREADS)
compact_kernel(const uint8_t* __restrict__ flags, int n, const float2* __restrict__ a_in, float2* __restrict__ a_out,
               const float2* __restrict__ b_in, float2* __restrict__ b_out, const float3* __restrict__ c_in,
               float3* __restrict__ c_out, int32_t* __restrict__ idx_out, int* __restrict__ count_out) {
  __shared__ int warp_tot[32];
  const int t = threadIdx.x;
  const int chunk = (n + COMPACT_THREADS - 1) / COMPACT_THREADS;
  const int beg = t * chunk;
  const int end = min(beg + chunk, n);
  int cnt = 0;
  for (int i = beg; i < end; i++) cnt += (flags[i] == 1);
  // block exclusive scan
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((t & 31) >= d) incl += v;
  }
  if ((t & 31) == 31) warp_tot[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    int w = warp_tot[t];
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, d);
      if (t >= d) wi += v;
    }
    warp_tot[t] = wi - w;  // exclusive
    if (t == 31) *count_out = wi;
  }
  __syncthreads();
  int pos = warp_tot[t >> 5] + incl - cnt;
  for (int i = beg; i < end; i++) {
    if (flags[i] == 1) {
      if (a_in) a_out[pos] = a_in[i];
      if (b_in) b_out[pos] = b_in[i];
      if (c_in) c_out[pos] = c_in[i];
      if (idx_out) idx_out[pos] = i;
      pos++;
    }
  }
}

int compact_launch(vo_ctx* c, const uint8_t* d_flags, int n, const float2* a_in, float2* a_out, const float2* b_in,
                   float2* b_out, const float3* c_in, float3* c_out, int32_t* idx_out, int count_slot) {
  if (n <= 0) {
    VO_CUDA(cudaMemsetAsync(c->d_count + count_slot, 0, sizeof(int), c->stream));
    return VO_OK;
  }
  {
    LaunchScope ls(c, VO_K_COMPACT);
    compact_kernel<<<1, COMPACT_THREADS, 0, c->stream>>>(d_flags, n, a_in, a_out, b_in, b_out, c_in, c_out, idx_out,
                                                         c->d_count + count_slot);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ---------------------------------------------------------------------------- triangulation
// One thread per correspondence: 4x4 DLT matrix in FP64, OpenCV's Jacobi SVD (bit-exact
// operation order, cvmath.cuh), last right-singular vector -> float -> divide by w in
// float.  Optional fused epilogue: the keyframe rigid transform p' = M * [p;1] in double.
__device__ __forceinline__ float3 rigid_apply(const double* M, float3 p) {
  const double x = p.x, y = p.y, z = p.z;
  float3 o;
  o.x = (float)(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[0], x), __dmul_rn(M[1], y)), __dmul_rn(M[2], z)), M[3]));
  o.y = (float)(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[4], x), __dmul_rn(M[5], y)), __dmul_rn(M[6], z)), M[7]));
  o.z = (float)(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[8], x), __dmul_rn(M[9], y)), __dmul_rn(M[10], z)), M[11]));
  return o;
}

__global__ void __launch_bounds__(128)
triangulate_kernel(const double* __restrict__ P, const float2* __restrict__ a, const float2* __restrict__ b, int n,
                   float3* __restrict__ out, const double* __restrict__ M, float3* __restrict__ out2) {
  __shared__ double sP[24], sM[12];
  if (threadIdx.x < 24) sP[threadIdx.x] = P[threadIdx.x];
  if (M && threadIdx.x < 12) sM[threadIdx.x] = M[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 p1 = a[i], p2 = b[i];
  float xyz[3];
  triangulate_dlt(sP, sP + 12, p1.x, p1.y, p2.x, p2.y, xyz);
  const float3 r = make_float3(xyz[0], xyz[1], xyz[2]);
  out[i] = r;
  if (M) out2[i] = rigid_apply(sM, r);
}

int triangulate_launch(vo_ctx* c, const double* d_P1P2, const float2* a, const float2* b, int n, float3* out,
                       const double* d_M, float3* out2) {
  if (n <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_TRIANGULATE);
    triangulate_kernel<<<div_up(n, 128), 128, 0, c->stream>>>(d_P1P2, a, b, n, out, d_M, out2);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

__global__ void transform_kernel(const double* __restrict__ M, const float3* __restrict__ in, int n,
                                 float3* __restrict__ out) {
  __shared__ double sM[12];
  if (threadIdx.x < 12) sM[threadIdx.x] = M[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = rigid_apply(sM, in[i]);
}

int transform_launch(vo_ctx* c, const double* d_M, const float3* in, int n, float3* out) {
  if (n <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_MISC);
    transform_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_M, in, n, out);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
