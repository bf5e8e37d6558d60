// selfcheck.cu -- device-vs-host bit comparison of the FP64 solver numerics.
//
// The RANSAC parity claim ("same inlier set for the same sample list") rests on the device
// executing cvmath.cuh / fmat7.cuh with exactly the rounding of an IEEE host.  During
// bring-up one build of this library produced EPnP results that differed from the host
// build of the same header although the SASS had the same FP64 instruction mix (a code
// shape-dependent front-end issue; see DESIGN.md "compiler sensitivity").  vo_create()
// therefore runs every solver on canned inputs on the device and on the host (this same
// translation unit compiled for x86) and refuses to create a context if a single bit
// differs.  The host results are only compared, never returned: this is a guard, not a
// CPU path.
#include "common.cuh"
#include "cvmath.cuh"
#include "fmat7.cuh"

namespace vo {

constexpr int SC_N = 24;        // canned cases per solver
constexpr int SC_EPNP = 12;     // doubles per EPnP result (R 9, t 3)
constexpr int SC_F = 28;        // n + 27 doubles per 7-point result
constexpr int SC_T = 3;         // floats per triangulation (stored as doubles)
constexpr int SC_STRIDE = SC_EPNP + SC_F + SC_T;

struct SelfCase {
  float obj[15], img[10];
  float m1[14], m2[14];
  float t1[2], t2[2];
};

__host__ __device__ static void selfcheck_eval(const SelfCase& c, const Intrinsics& K, const double* P, double* out) {
  double R[9], t[3];
  epnp5<false>(c.obj, c.img, K, R, t);
  for (int i = 0; i < 9; i++) out[i] = R[i];
  for (int i = 0; i < 3; i++) out[9 + i] = t[i];
  double F[27];
  for (int i = 0; i < 27; i++) F[i] = 0;
  const int n = fmat_7point(c.m1, c.m2, F);
  out[SC_EPNP] = n;
  for (int i = 0; i < 27; i++) out[SC_EPNP + 1 + i] = (i < 9 * (n > 0 ? n : 0)) ? F[i] : 0;
  float xyz[3];
  triangulate_dlt(P, P + 12, c.t1[0], c.t1[1], c.t2[0], c.t2[1], xyz);
  for (int i = 0; i < 3; i++) out[SC_EPNP + SC_F + i] = xyz[i];
}

__global__ void selfcheck_kernel(const SelfCase* cases, int n, Intrinsics K, const double* P, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  selfcheck_eval(cases[i], K, P, out + (size_t)i * SC_STRIDE);
}

int selfcheck_run(vo_ctx* c) {
#ifdef VO_NO_SELFCHECK   // instrumented debug builds only (tools/); never defined by build.py
  (void)c;
  return VO_OK;
#endif
  const Intrinsics K{c->p.fx, c->p.fy, c->p.cx, c->p.cy};
  double P[24];
  {
    const double Km[9] = {K.fx, 0, K.cx, 0, K.fy, K.cy, 0, 0, 1};
    const double E1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const double E2[12] = {1, 0, 0, -c->p.baseline, 0, 1, 0, 0, 0, 0, 1, 0};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 4; j++) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 3; k++) {
          s1 += Km[i * 3 + k] * E1[k * 4 + j];
          s2 += Km[i * 3 + k] * E2[k * 4 + j];
        }
        P[i * 4 + j] = s1;
        P[12 + i * 4 + j] = s2;
      }
  }
  std::vector<SelfCase> cases(SC_N);
  CvRng rng(0x9e3779b97f4a7c15ULL);
  auto uni = [&](double a, double b) { return a + (b - a) * (rng.next() * (1.0 / 4294967296.0)); };
  for (int s = 0; s < SC_N; s++) {
    SelfCase& cs = cases[s];
    for (int i = 0; i < 5; i++) {
      const double x = uni(-20, 20), y = uni(-3, 3), z = uni(4, 60);
      cs.obj[3 * i] = (float)x; cs.obj[3 * i + 1] = (float)y; cs.obj[3 * i + 2] = (float)z;
      const double zc = z - 0.8;
      cs.img[2 * i] = (float)(K.fx * (x + 0.05) / zc + K.cx + uni(-0.5, 0.5) + (s % 4 == 3 && i == 2 ? 40.0 : 0.0));
      cs.img[2 * i + 1] = (float)(K.fy * (y - 0.02) / zc + K.cy + uni(-0.5, 0.5));
    }
    for (int i = 0; i < 7; i++) {
      const double u = uni(0, c->p.width), v = uni(0, c->p.height), z = uni(4, 60);
      cs.m1[2 * i] = (float)u; cs.m1[2 * i + 1] = (float)v;
      cs.m2[2 * i] = (float)(u + (u - K.cx) * 0.85 / z + uni(-0.3, 0.3));
      cs.m2[2 * i + 1] = (float)(v + (v - K.cy) * 0.85 / z + uni(-0.3, 0.3));
    }
    const double u = uni(0, c->p.width), v = uni(0, c->p.height), z = uni(3, 80);
    cs.t1[0] = (float)u; cs.t1[1] = (float)v;
    cs.t2[0] = (float)(u - K.fx * c->p.baseline / z); cs.t2[1] = (float)(v + uni(-0.1, 0.1));
  }
  std::vector<double> host((size_t)SC_N * SC_STRIDE), dev((size_t)SC_N * SC_STRIDE);
  for (int s = 0; s < SC_N; s++) selfcheck_eval(cases[s], K, P, host.data() + (size_t)s * SC_STRIDE);

  SelfCase* d_cases = nullptr;
  double *d_P = nullptr, *d_out = nullptr;
  VO_CUDA(cudaMalloc(&d_cases, SC_N * sizeof(SelfCase)));
  VO_CUDA(cudaMalloc(&d_P, sizeof(P)));
  VO_CUDA(cudaMalloc(&d_out, dev.size() * sizeof(double)));
  VO_CUDA(cudaMemcpyAsync(d_cases, cases.data(), SC_N * sizeof(SelfCase), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(d_P, P, sizeof(P), cudaMemcpyHostToDevice, c->stream));
  c->launch_count++;
  selfcheck_kernel<<<div_up(SC_N, 8), 8, 0, c->stream>>>(d_cases, SC_N, K, d_P, d_out);
  VO_CUDA(cudaGetLastError());
  VO_CUDA(cudaMemcpyAsync(dev.data(), d_out, dev.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_cases);
  cudaFree(d_P);
  cudaFree(d_out);
  // ---- the production RANSAC solve kernels (warp-cooperative Jacobi) on the same cases
  int bad_prod = 0;
  double worst_prod_f = 0;
  {
    std::vector<float> xyz(SC_N * 15), xy(SC_N * 10), f1(SC_N * 14), f2(SC_N * 14);
    std::vector<int32_t> s5(SC_N * 5), s7(SC_N * 7);
    for (int s = 0; s < SC_N; s++) {
      memcpy(&xyz[s * 15], cases[s].obj, 15 * sizeof(float));
      memcpy(&xy[s * 10], cases[s].img, 10 * sizeof(float));
      memcpy(&f1[s * 14], cases[s].m1, 14 * sizeof(float));
      memcpy(&f2[s * 14], cases[s].m2, 14 * sizeof(float));
      for (int i = 0; i < 5; i++) s5[s * 5 + i] = s * 5 + i;
      for (int i = 0; i < 7; i++) s7[s * 7 + i] = s * 7 + i;
    }
    VO_CUDA(cudaMemcpyAsync(c->d_f_xyz, xyz.data(), xyz.size() * 4, cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->d_f_trk, xy.data(), xy.size() * 4, cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->d_samples, s5.data(), s5.size() * 4, cudaMemcpyHostToDevice, c->stream));
    VO_TRY(pnp_solve_launch(c, c->d_f_xyz, c->d_f_trk, c->d_samples, SC_N, c->d_models, c->d_counts));
    std::vector<double> pm((size_t)SC_N * 16);
    VO_CUDA(cudaMemcpyAsync(pm.data(), c->d_models, pm.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < SC_N; s++) {
      // tvec has no libm on its path: bit comparison.  rvec goes through acos/sin/cos.
      const double* h = host.data() + (size_t)s * SC_STRIDE;
      if (memcmp(h + 9, &pm[(size_t)s * 16 + 3], 3 * sizeof(double)) != 0) bad_prod++;
      double rv[3];
      rodrigues_mat2vec(h, rv);
      for (int i = 0; i < 3; i++)
        if (fabs(rv[i] - pm[(size_t)s * 16 + i]) > 1e-12) bad_prod++;
    }
    VO_CUDA(cudaMemcpyAsync(c->d_c_ref, f1.data(), f1.size() * 4, cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->d_c_trk, f2.data(), f2.size() * 4, cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->d_samples, s7.data(), s7.size() * 4, cudaMemcpyHostToDevice, c->stream));
    VO_TRY(fmat_solve_launch(c, c->d_c_ref, c->d_c_trk, c->d_samples, SC_N, c->d_models, c->d_counts));
    std::vector<double> fm((size_t)SC_N * 27);
    std::vector<int32_t> fc((size_t)SC_N * 3);
    VO_CUDA(cudaMemcpyAsync(fm.data(), c->d_models, fm.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(fc.data(), c->d_counts, fc.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < SC_N; s++) {
      const double* h = host.data() + (size_t)s * SC_STRIDE + SC_EPNP;
      const int n = (int)h[0];
      for (int k = 0; k < 3; k++) {
        if ((fc[s * 3 + k] >= 0) != (k < n)) bad_prod++;
        if (k < n)
          for (int i = 0; i < 9; i++) {
            const double a = h[1 + k * 9 + i], b = fm[(size_t)(s * 3 + k) * 9 + i];
            const double e = fabs(a - b) / fmax(fabs(a), 1e-12);
            if (e > worst_prod_f) worst_prod_f = e;
          }
      }
    }
  }
  int bad_epnp = 0, bad_tri = 0, bad_f = 0;
  double worst_f = 0;
  for (int s = 0; s < SC_N; s++) {
    const double* h = host.data() + (size_t)s * SC_STRIDE;
    const double* d = dev.data() + (size_t)s * SC_STRIDE;
    if (memcmp(h, d, SC_EPNP * sizeof(double)) != 0) bad_epnp++;
    if (memcmp(h + SC_EPNP + SC_F, d + SC_EPNP + SC_F, SC_T * sizeof(double)) != 0) bad_tri++;
    // the 7-point solver goes through acos/cos/pow (cubic roots): device libm differs from
    // glibc in the last bits, so F is compared to 1e-9 relative and the model count exactly
    if (h[SC_EPNP] != d[SC_EPNP]) bad_f++;
    for (int i = 0; i < 27; i++) {
      const double a = h[SC_EPNP + 1 + i], b = d[SC_EPNP + 1 + i];
      const double e = fabs(a - b) / fmax(fabs(a), 1e-12);
      if (e > worst_f) worst_f = e;
    }
  }
  if (bad_epnp || bad_tri || bad_f || worst_f > 1e-9 || bad_prod || worst_prod_f > 1e-9) {
    set_error("numerics self-check failed: device != host for %d/%d EPnP, %d/%d triangulation, %d/%d 7-point cases "
              "(F rel err %.3g), %d mismatches in the production solve kernels (F rel err %.3g); this build of "
              "libvo_b200 must not be used",
              bad_epnp, SC_N, bad_tri, SC_N, bad_f, SC_N, worst_f, bad_prod, worst_prod_f);
    return VO_ERR_SELF_CHECK;
  }
  return VO_OK;
}

}  // namespace vo
