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
// Order-preserving stream compaction of up to three parallel arrays (+ the index list), single
// pass with decoupled look-back: every CTA owns a tile of CP_TILE flags (8 per thread, one
// 64-bit load), scans it, publishes its tile total tagged with the launch epoch, and sums the
// totals of the lower-indexed tiles (at most 64 tiles, all co-resident, so the wait is short
// and cannot deadlock).  The arrays are L2-resident (<= 1.5 MB): latency-, not bandwidth-bound.
__global__ void __launch_bounds__(CP_THREADS)
compact_kernel(const uint8_t* __restrict__ flags, int n, const float2* __restrict__ a_in, float2* __restrict__ a_out,
               const float2* __restrict__ b_in, float2* __restrict__ b_out, const float3* __restrict__ c_in,
               float3* __restrict__ c_out, int32_t* __restrict__ idx_out, int* __restrict__ count_out,
               volatile unsigned long long* tile_state, unsigned* epoch_ctr, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const int t = threadIdx.x;
  const int beg = blockIdx.x * CP_TILE + t * CP_ITEMS;
  unsigned long long bits = 0;  // byte k = flag of element beg+k
  if (beg + CP_ITEMS <= n) {
    bits = *reinterpret_cast<const unsigned long long*>(flags + beg);
  } else {
    for (int k = 0; k < CP_ITEMS; k++)
      if (beg + k < n) bits |= (unsigned long long)flags[beg + k] << (8 * k);
  }
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; k++) cnt += ((bits >> (8 * k)) & 0xff) == 1;
  int pos = compact_tile_offset(cnt, count_out, tile_state, epoch_ctr);
#pragma unroll
  for (int k = 0; k < CP_ITEMS; k++) {
    if (((bits >> (8 * k)) & 0xff) == 1) {
      const int i = beg + k;
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
    compact_kernel<<<div_up(n, CP_TILE), CP_THREADS, 0, c->stream>>>(d_flags, n, a_in, a_out, b_in, b_out, c_in, c_out,
                                                                     idx_out, c->d_count + count_slot, c->d_tile_state,
                                                                     c->d_epoch, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ---------------------------------------------------------------------------- look-ahead gather
__global__ void __launch_bounds__(256)
gather_tracks_kernel(const int32_t* __restrict__ idx, int n, const float2* __restrict__ trk_in, const uint8_t* __restrict__ st_in,
                     float2* __restrict__ trk_out, uint8_t* __restrict__ st_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = idx[i];
  trk_out[i] = trk_in[j];
  st_out[i] = st_in[j];
}

int gather_tracks_launch(vo_ctx* c, const int32_t* d_idx, int n, const float2* trk_in, const uint8_t* st_in, float2* trk_out,
                         uint8_t* st_out) {
  if (n <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_COMPACT);
    gather_tracks_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_idx, n, trk_in, st_in, trk_out, st_out);
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
                   float3* __restrict__ out, const double* __restrict__ M, float3* __restrict__ out2,
                   const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
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
    triangulate_kernel<<<div_up(n, 128), 128, 0, c->stream>>>(d_P1P2, a, b, n, out, d_M, out2, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

__global__ void transform_kernel(const double* __restrict__ M, const float3* __restrict__ in, int n,
                                 float3* __restrict__ out, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  __shared__ double sM[12];
  if (threadIdx.x < 12) sM[threadIdx.x] = M[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = rigid_apply(sM, in[i]);
}

// The keyframe epilogue of the sequence driver in one launch: world points = pose * camera points (the pose travels as a
// kernel argument, no upload) and the keyframe's 2-D points copied next to them.
struct Pose3x4 {
  double m[12];
};

__global__ void keyframe_epilogue_kernel(const Pose3x4 M, const float3* __restrict__ cam, const float2* __restrict__ xy_in, int n,
                                         float3* __restrict__ world, float2* __restrict__ xy_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  world[i] = rigid_apply(M.m, cam[i]);
  xy_out[i] = xy_in[i];
}

int keyframe_epilogue_launch(vo_ctx* c, const double* pose3x4, const float3* cam, const float2* xy_in, int n, float3* world,
                             float2* xy_out) {
  if (n <= 0) return VO_OK;
  Pose3x4 M;
  for (int i = 0; i < 12; i++) M.m[i] = pose3x4[i];
  {
    LaunchScope ls(c, VO_K_MISC);
    keyframe_epilogue_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(M, cam, xy_in, n, world, xy_out);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int transform_launch(vo_ctx* c, const double* d_M, const float3* in, int n, float3* out) {
  if (n <= 0) return VO_OK;
  {
    LaunchScope ls(c, VO_K_MISC);
    transform_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_M, in, n, out, c->n_dev);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

}  // namespace vo
