// TMA bring-up probe (development tool): which box / coordinate / extent combinations does a 2-D u8
// cp.async.bulk.tensor load accept?   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

template <int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, unsigned* out) {
  __shared__ alignas(128) uint8_t t0[BW * BH];
  __shared__ alignas(8) unsigned long long bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(BW * BH) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(t0)),
                 "l"(reinterpret_cast<uint64_t>(&tm)), "r"(x), "r"(y), "r"(bar_a)
                 : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(bar_a), "r"(0)
      : "memory");
  unsigned s = 0;
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) s += t0[i];
  atomicAdd(out, s);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BW, int BH>
int run(EncodeFn enc, uint8_t* d, int w, int h, int pitch, int x, int y) {
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {BW, BH};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  unsigned* out;
  cudaMalloc(&out, 4);
  cudaMemset(out, 0, 4);
  k<BW, BH><<<1, 256>>>(tm, x, y, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned v = 0;
  cudaMemcpy(&v, out, 4, cudaMemcpyDeviceToHost);
  printf("box %dx%d dims %dx%d pitch %d at (%d,%d): %s sum=%u\n", BW, BH, w, h, pitch, x, y, cudaGetErrorString(e), v);
  return e != cudaSuccess;
}

int main(int argc, char** argv) {
  int test = argc > 1 ? atoi(argv[1]) : 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (!fn) { printf("no entry point\n"); return 2; }
  EncodeFn enc = (EncodeFn)fn;
  const int pitch = 1408, rows = 376 + 42;
  uint8_t* d;
  cudaMalloc(&d, (size_t)pitch * rows);
  cudaMemset(d, 1, (size_t)pitch * rows);
  uint8_t* base = d + 21 * pitch + 32;
  switch (test) {
    case 0: return run<128, 64>(enc, base, 1280, 376, pitch, 0, 0);
    case 1: return run<128, 64>(enc, base, 1241, 376, pitch, 0, 0);
    case 2: return run<112, 69>(enc, base, 1241, 376, pitch, 0, 0);
    case 3: return run<112, 69>(enc, base, 1241, 376, pitch, 42, 10);
    case 4: return run<112, 69>(enc, base, 1241, 376, pitch, 1194, 330);
    case 5: return run<128, 64>(enc, base, 1241, 376, pitch, 3, 5);
    case 6: return run<112, 64>(enc, base, 1241, 376, pitch, 0, 0);
    case 7: return run<128, 69>(enc, base, 1241, 376, pitch, 0, 0);
  }
  return 0;
}
