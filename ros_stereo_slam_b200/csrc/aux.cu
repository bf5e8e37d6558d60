// aux.cu -- K9b ANMS, plus harness-only kernels (synthetic scene renderer, FP32 issue-rate
// microbenchmark).
//
// ANMS replaces adaptiveNonMaximalSuppresion (reference src/ANMS.cpp:18-67): sort by response
// (descending; ties by original index -- std::sort there is unstable, only the kept SET is
// defined), suppression radius = distance to the nearest keypoint whose response exceeds
// 1.11x ours (float product, double distance, as in ANMS.cpp:41-51), keep everything whose
// radius >= the (numToKeep+1)-th largest radius.
#include <float.h>
#include <math.h>

#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"

namespace vo {

// ------------------------------------------------------------------------------------ ANMS
__global__ void anms_gather_kernel(const float2* __restrict__ xy, const int* __restrict__ order, int n,
                                   float2* __restrict__ sxy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sxy[i] = xy[order[i]];
}

constexpr int ANMS_TPB = 256;

// resp: sorted descending.  radius2[i] = min_{j < min(i, stop_i)} |p_i - p_j|^2 (double), where
// stop_i = first j with !(resp[j] > resp[i]*1.11f); DBL_MAX if the range is empty.
__global__ void __launch_bounds__(ANMS_TPB)
anms_radius_kernel(const float2* __restrict__ sxy, const float* __restrict__ resp, int n, double* __restrict__ radius) {
  __shared__ float2 tile[ANMS_TPB];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float2 p = make_float2(0, 0);
  int bound = 0;
  if (i < n) {
    p = sxy[i];
    const float r = __fmul_rn(resp[i], 1.11f);
    int lo = 0, hi = i;  // first j in [0, i) with !(resp[j] > r)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (resp[mid] > r) lo = mid + 1; else hi = mid;
    }
    bound = lo;
  }
  // block-wide max bound
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  atomicMax(&s_max, bound);
  __syncthreads();
  const int jmax = s_max;
  double best = DBL_MAX;
  for (int j0 = 0; j0 < jmax; j0 += ANMS_TPB) {
    const int j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < n ? sxy[j] : make_float2(0, 0);
    __syncthreads();
    const int lim = min(ANMS_TPB, bound - j0);
    for (int k = 0; k < lim; k++) {
      const float dx = __fsub_rn(p.x, tile[k].x), dy = __fsub_rn(p.y, tile[k].y);
      const double d2 = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy));
      best = d2 < best ? d2 : best;
    }
    __syncthreads();
  }
  if (i < n) radius[i] = best == DBL_MAX ? DBL_MAX : sqrt(best);
}

__global__ void anms_flag_kernel(const double* __restrict__ radius, const double* __restrict__ sorted_desc, int n,
                                 int num_keep, uint8_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = radius[i] >= sorted_desc[num_keep];
}

int anms_launch(vo_ctx* c, const float* h_xy, const float* h_resp, int n, int num_keep, int32_t* keep_idx, int cap,
                int* n_keep) {
  *n_keep = 0;
  if (n > c->cap) return VO_ERR_CAPACITY;
  if (n < num_keep) {  // ANMS.cpp:21 -- nothing to do, everything is kept (input order)
    if (n > cap) return VO_ERR_CAPACITY;
    for (int i = 0; i < n; i++) keep_idx[i] = i;
    *n_keep = n;
    return VO_OK;
  }
  if (n == num_keep) {
    set_error("ANMS: size == numToKeep reads radiiSorted[numToKeep] out of bounds in the reference (ANMS.cpp:59)");
    return VO_ERR_INVALID_ARG;
  }
  if (n == 0) return VO_OK;
  // scratch
  float *d_resp_in = nullptr, *d_resp = nullptr;
  int *d_ord_in = nullptr, *d_ord = nullptr;
  double *d_rad = nullptr, *d_rad_sorted = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0, tb2 = 0;
  int rc = VO_OK;
  auto fail = [&](cudaError_t e) {
    set_error("anms: %s", cudaGetErrorString(e));
    rc = VO_ERR_CUDA;
  };
#define ANMS_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { fail(_e); goto done; } } while (0)
  {
    ANMS_CUDA(cudaMalloc(&d_resp_in, n * sizeof(float)));
    ANMS_CUDA(cudaMalloc(&d_resp, n * sizeof(float)));
    ANMS_CUDA(cudaMalloc(&d_ord_in, n * sizeof(int)));
    ANMS_CUDA(cudaMalloc(&d_ord, n * sizeof(int)));
    ANMS_CUDA(cudaMalloc(&d_rad, n * sizeof(double)));
    ANMS_CUDA(cudaMalloc(&d_rad_sorted, n * sizeof(double)));
    std::vector<int> iota(n);
    for (int i = 0; i < n; i++) iota[i] = i;
    ANMS_CUDA(cudaMemcpyAsync(d_resp_in, h_resp, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    ANMS_CUDA(cudaMemcpyAsync(d_ord_in, iota.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    ANMS_CUDA(cudaMemcpyAsync(c->d_xy_in, h_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    ANMS_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_resp_in, d_resp, d_ord_in, d_ord, n, 0, 32,
                                                        c->stream));
    ANMS_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, tb2, d_rad, d_rad_sorted, n, 0, 64, c->stream));
    tmp_bytes = std::max(tmp_bytes, tb2);
    ANMS_CUDA(cudaMalloc(&d_tmp, tmp_bytes));
    c->launch_count += 2;
    ANMS_CUDA(cub::DeviceRadixSort::SortPairsDescending(d_tmp, tmp_bytes, d_resp_in, d_resp, d_ord_in, d_ord, n, 0, 32,
                                                        c->stream));  // stable: ties keep index order
    {
      LaunchScope ls(c, VO_K_MISC);
      anms_gather_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(c->d_xy_in, d_ord, n, c->d_xy_trk);
    }
    {
      LaunchScope ls(c, VO_K_MISC);
      anms_radius_kernel<<<div_up(n, ANMS_TPB), ANMS_TPB, 0, c->stream>>>(c->d_xy_trk, d_resp, n, d_rad);
    }
    ANMS_CUDA(cub::DeviceRadixSort::SortKeysDescending(d_tmp, tmp_bytes, d_rad, d_rad_sorted, n, 0, 64, c->stream));
    {
      LaunchScope ls(c, VO_K_MISC);
      anms_flag_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_rad, d_rad_sorted, n, num_keep, c->d_mask);
    }
    // ordered compaction of the ORIGINAL indices (d_ord) by flag
    rc = compact_launch(c, c->d_mask, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c->d_idx, 5);
    if (rc != VO_OK) goto done;
    ANMS_CUDA(cudaMemcpyAsync(c->h_count, c->d_count, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ANMS_CUDA(cudaStreamSynchronize(c->stream));
    const int k = c->h_count[5];
    *n_keep = k;
    if (k > cap) {
      rc = VO_ERR_CAPACITY;
      goto done;
    }
    // d_idx holds positions in sorted order; map through d_ord on the host side
    std::vector<int> pos(k), ord(n);
    ANMS_CUDA(cudaMemcpyAsync(pos.data(), c->d_idx, k * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ANMS_CUDA(cudaMemcpyAsync(ord.data(), d_ord, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ANMS_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < k; i++) keep_idx[i] = ord[pos[i]];
  }
done:
  cudaFree(d_resp_in); cudaFree(d_resp); cudaFree(d_ord_in); cudaFree(d_ord);
  cudaFree(d_rad); cudaFree(d_rad_sorted); cudaFree(d_tmp);
  return rc;
#undef ANMS_CUDA
}

// ------------------------------------------------------------------------------------ synthetic scene (harness)
// Same scene as oracle/synth.py (ground, two walls, far ceiling, hashed boxes; multi-octave
// value-noise texture with anti-alias fade and fog), evaluated in double per pixel.
namespace synth {
constexpr double FX = 718.856, FY = 718.856, CX = 607.1928, CY = 185.2157, BASE = 0.54;
constexpr double GROUND_Y = 1.65, WALL_L = -7.5, WALL_R = 8.5, CELL = 8.0, FOG = 160.0;
constexpr int NBOX = 20;

struct Box { double x0, x1, y0, y1, z0, z1; long long id; int valid; };
struct Frame {
  double R[9];
  double c[3];
  Box box[NBOX];
  int seed;
};

__host__ __device__ inline uint32_t hash_u32(long long ix, long long iy, long long seed) {
  const uint32_t x = (uint32_t)(uint64_t)ix, y = (uint32_t)(uint64_t)iy, s = (uint32_t)(uint64_t)seed;
  uint32_t h = x * 374761393u + y * 668265263u + s * 2246822519u;
  h = (h ^ (h >> 13)) * 1274126177u;
  h = h ^ (h >> 16);
  return h;
}
__host__ __device__ inline double hash01(long long ix, long long iy, long long seed) {
  return (double)(hash_u32(ix, iy, seed) >> 8) * (1.0 / 16777216.0);
}

__host__ __device__ inline long long pmod(long long a, long long m) { return ((a % m) + m) % m; }  // Python's %

__device__ inline double value_noise(double u, double v, long long seed) {
  const double fu = floor(u), fv = floor(v);
  const long long iu = (long long)fu, iv = (long long)fv;
  double a = u - fu, b = v - fv;
  a = a * a * (3 - 2 * a);
  b = b * b * (3 - 2 * b);
  const double n00 = hash01(iu, iv, seed), n10 = hash01(iu + 1, iv, seed);
  const double n01 = hash01(iu, iv + 1, seed), n11 = hash01(iu + 1, iv + 1, seed);
  const double n0 = n00 + (n10 - n00) * a, n1 = n01 + (n11 - n01) * a;
  return 2.0 * (n0 + (n1 - n0) * b) - 1.0;
}

__device__ inline double texture(double u, double v, double footprint, long long seed) {
  const double freq[4] = {0.9, 2.3, 6.1, 17.0}, amp[4] = {1.0, 0.8, 0.65, 0.5};
  double acc = 0, wsum = 0;
  for (int k = 0; k < 4; k++) {
    const double period_px = 1.0 / (freq[k] * fmax(footprint, 1e-9));
    const double fade = fmin(fmax((period_px - 2.0) * 0.5, 0.0), 1.0);
    acc += amp[k] * fade * value_noise(u * freq[k], v * freq[k], seed * 4 + k);
    wsum += amp[k];
  }
  return acc / wsum;
}

__global__ void render_kernel(Frame f, int width, int height, uint8_t* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= width || v >= height) return;
  const double dcx = (u - CX) / FX, dcy = (v - CY) / FY, dcz = 1.0;
  const double dx = f.R[0] * dcx + f.R[1] * dcy + f.R[2] * dcz;
  const double dy = f.R[3] * dcx + f.R[4] * dcy + f.R[5] * dcz;
  const double dz = f.R[6] * dcx + f.R[7] * dcy + f.R[8] * dcz;
  const double ox = f.c[0], oy = f.c[1], oz = f.c[2];
  double tb = 1e30, tu = 0, tv = 0, cn = 1;
  long long sid = 0;
  auto consider = [&](double t, double a, double b, double cc, long long id, bool valid) {
    if (valid && t > 1e-3 && t < tb) { tb = t; tu = a; tv = b; cn = cc; sid = id; }
  };
  {
    double t = (GROUND_Y - oy) / dy;
    consider(t, ox + t * dx, oz + t * dz, fabs(dy), 1, dy > 1e-9);
    t = (WALL_L - ox) / dx;
    consider(t, oz + t * dz, oy + t * dy, fabs(dx), 2, dx < -1e-9);
    t = (WALL_R - ox) / dx;
    consider(t, oz + t * dz, oy + t * dy, fabs(dx), 3, dx > 1e-9);
    t = (-12.0 - oy) / dy;
    consider(t, ox + t * dx, oz + t * dz, fabs(dy), 4, dy < -1e-9);
  }
  for (int b = 0; b < NBOX; b++) {
    const Box& bx = f.box[b];
    if (!bx.valid) continue;
    const double tx0 = (bx.x0 - ox) / dx, tx1 = (bx.x1 - ox) / dx;
    const double ty0 = (bx.y0 - oy) / dy, ty1 = (bx.y1 - oy) / dy;
    const double tz0 = (bx.z0 - oz) / dz, tz1 = (bx.z1 - oz) / dz;
    const double tnx = fmin(tx0, tx1), tfx = fmax(tx0, tx1);
    const double tny = fmin(ty0, ty1), tfy = fmax(ty0, ty1);
    const double tnz = fmin(tz0, tz1), tfz = fmax(tz0, tz1);
    const double tn = fmax(fmax(tnx, tny), tnz), tf = fmin(fmin(tfx, tfy), tfz);
    const bool hit = (tn <= tf) && (tn > 1e-3);
    const double px = ox + tn * dx, py = oy + tn * dy, pz = oz + tn * dz;
    const bool face_x = (tnx >= tny) && (tnx >= tnz);
    const bool face_y = (!face_x) && (tny >= tnz);
    const double a = (face_x ? pz : px) + 13.7 * (double)pmod(bx.id, 97);
    const double bb = (face_y ? pz : py) + 7.3 * (double)pmod(bx.id, 89);
    const double cc = face_x ? fabs(dx) : (face_y ? fabs(dy) : fabs(dz));
    consider(tn, a, bb, cc, 5 + pmod(bx.id, 1000) * 3 + (face_x ? 0 : (face_y ? 1 : 2)), hit);
  }
  const double tt = tb < 1e30 ? tb : 1e4;
  const double dn = sqrt(dx * dx + dy * dy + dz * dz);
  const double footprint = tt * dn / FX / fmax(cn / dn, 0.05);
  const double val = texture(tu, tv, footprint, (long long)f.seed * 131 + sid);
  const double fog = exp(-(tt * dn) / FOG);
  double img = rint(128.0 + 118.0 * val * fog);
  img = fmin(fmax(img, 0.0), 255.0);
  out[(size_t)v * width + u] = (uint8_t)img;
}
}  // namespace synth

int synth_launch(vo_ctx* c, int seed, int frame, int eye, uint8_t* d_out) {
  using namespace synth;
  const double speed = 0.85, yaw_amp = 4.0 * M_PI / 180.0, yaw_rate = 0.1;
  Frame f;
  f.seed = seed;
  double x = 0, z = 0;
  for (int k = 0; k < frame; k++) {
    const double y = yaw_amp * sin(yaw_rate * k);
    x += speed * sin(y);
    z += speed * cos(y);
  }
  const double yw = yaw_amp * sin(yaw_rate * frame);
  const double cs = cos(yw), sn = sin(yw);
  const double R[9] = {cs, 0, sn, 0, 1, 0, -sn, 0, cs};
  memcpy(f.R, R, sizeof(R));
  f.c[0] = x; f.c[1] = 0; f.c[2] = z;
  if (eye == 1) {
    f.c[0] += R[0] * BASE;
    f.c[1] += R[3] * BASE;
    f.c[2] += R[6] * BASE;
  }
  const long long c0 = (long long)floor(f.c[2] / CELL) - 1;
  for (int b = 0; b < NBOX; b++) {
    const long long cell = c0 + b;
    double r[6];
    for (int k = 0; k < 6; k++) r[k] = hash01(cell, k, (long long)seed + 101);
    Box& bx = f.box[b];
    bx.valid = r[0] <= 0.8;
    const double w = 0.8 + 1.7 * r[1], h = 0.8 + 2.2 * r[2], d = 0.8 + 2.2 * r[3];
    const double xc = r[4] < 0.5 ? -6.0 + 3.0 * r[5] : 3.0 + 4.0 * r[5];
    const double zc = ((double)cell + 0.5) * CELL;
    bx.x0 = xc - w / 2; bx.x1 = xc + w / 2;
    bx.y0 = GROUND_Y - h; bx.y1 = GROUND_Y;
    bx.z0 = zc - d / 2; bx.z1 = zc + d / 2;
    bx.id = cell;
  }
  dim3 b(32, 8), g(div_up(c->p.width, 32), div_up(c->p.height, 8));
  {
    LaunchScope ls(c, VO_K_MISC);
    render_kernel<<<g, b, 0, c->stream>>>(f, c->p.width, c->p.height, d_out);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ------------------------------------------------------------------------------------ FP32 issue-rate peak
__global__ void fp32_peak_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float m = 0.999f, b = 1e-3f + blockIdx.x * 1e-9f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
      a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int fp32_peak_launch(vo_ctx* c, double* tflops) {
  const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
  float* d = nullptr;
  VO_CUDA(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t a, b;
  VO_CUDA(cudaEventCreate(&a));
  VO_CUDA(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    VO_CUDA(cudaEventRecord(a, c->stream));
    c->launch_count++;
    fp32_peak_kernel<<<blocks, threads, 0, c->stream>>>(d, iters);
    VO_CUDA(cudaEventRecord(b, c->stream));
    VO_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    VO_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  const double flops = (double)blocks * threads * iters * 16.0 * 8.0 * 2.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  return VO_OK;
}

// ------------------------------------------------------------------------------------ INT32 issue-rate peak
// The LK kernel is integer work (IMAD, DP2A, shifts): its roofline denominator is the IMAD issue rate, measured
// the same way as the FFMA one above (8 independent dependent chains per thread).
__global__ void int32_peak_kernel(int* out, int iters) {
  int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const int m = 3 + (int)blockIdx.x, b = 7 + (int)threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      a0 = a0 * m + b; a1 = a1 * m + b; a2 = a2 * m + b; a3 = a3 * m + b;
      a4 = a4 * m + b; a5 = a5 * m + b; a6 = a6 * m + b; a7 = a7 * m + b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int int32_peak_launch(vo_ctx* c, double* tops) {
  const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
  int* d = nullptr;
  VO_CUDA(cudaMalloc(&d, (size_t)blocks * threads * sizeof(int)));
  cudaEvent_t a, b;
  VO_CUDA(cudaEventCreate(&a));
  VO_CUDA(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    VO_CUDA(cudaEventRecord(a, c->stream));
    c->launch_count++;
    int32_peak_kernel<<<blocks, threads, 0, c->stream>>>(d, iters);
    VO_CUDA(cudaEventRecord(b, c->stream));
    VO_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    VO_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  const double ops = (double)blocks * threads * iters * 16.0 * 8.0 * 2.0;
  *tops = ops / (best * 1e-3) / 1e12;
  return VO_OK;
}

}  // namespace vo
