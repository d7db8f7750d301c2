// development aid: which fp64 TMA box shapes does the B200 accept?  nvcc -arch=sm_100a tma_box_test.cu -o tma_box_test -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void k(const __grid_constant__ CUtensorMap map, int bytes, int c0, int c1, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned bar = (unsigned)__cvta_generic_to_shared(smem + 8192);
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(&map), "r"(c0), "r"(c1), "r"(bar) : "memory");
  }
  __syncthreads();
  unsigned ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar) : "memory");
  if (threadIdx.x < 8) out[threadIdx.x] = ((double*)smem)[threadIdx.x];
}

struct Maps6 { CUtensorMap a, b, c, d, e, f; };
__global__ void k6(int pad0, int pad1, const __grid_constant__ Maps6 maps, double* out, int ba, int bf) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned bar = (unsigned)__cvta_generic_to_shared(smem + 12288);
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(ba + bf) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst + 10368), "l"(&maps.f), "r"(-5), "r"(-5), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(&maps.a), "r"(-6), "r"(-8), "r"(bar) : "memory");
  }
  __syncthreads();
  unsigned ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar) : "memory");
  if (threadIdx.x < 8) out[threadIdx.x] = ((double*)smem)[threadIdx.x];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  Enc enc = (Enc)ptr;
  const int rows = 40, cols = 31, ld = 32;
  double* d; cudaMalloc(&d, rows * ld * 8);
  double* h = (double*)malloc(rows * ld * 8);
  for (int i = 0; i < rows * ld; ++i) h[i] = i;
  cudaMemcpy(d, h, rows * ld * 8, cudaMemcpyHostToDevice);
  double* out; cudaMalloc(&out, 64);
  int shapes[][2] = {{64, 4}, {34, 3}, {36, 3}, {34, 4}, {34, 2}, {32, 3}, {66, 4}, {48, 3}};
  for (auto& sh : shapes) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)sh[0], (cuuint32_t)sh[1]};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box %dx%d: encode failed %d\n", sh[0], sh[1], (int)r); continue; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k<<<1, 32, 16384>>>(m, sh[0] * sh[1] * 8, -4, -1, out);
    cudaError_t e = cudaDeviceSynchronize();
    double o[8] = {0};
    if (e == cudaSuccess) cudaMemcpy(o, out, 64, cudaMemcpyDeviceToHost);
    printf("box %dx%d: %s  first row of box: %g %g %g %g %g %g\n", sh[0], sh[1], cudaGetErrorString(e), o[0], o[1], o[2], o[3], o[4], o[5]);
    if (e != cudaSuccess) { printf("(context lost, stopping)\n"); return 0; }
  }
  {
    Maps6 M;
    CUtensorMap ma, mf;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
    cuuint32_t es[2] = {1, 1};
    cuuint32_t boxa[2] = {64, 4}, boxf[2] = {34, 3};
    enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, boxa, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&mf, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, boxf, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M.a = M.b = M.c = M.d = M.e = ma; M.f = mf;
    cudaFuncSetAttribute(k6, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k6<<<1, 32, 16384>>>(0, 0, M, out, 64 * 4 * 8, 34 * 3 * 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("6-map struct, two boxes on one barrier, all-OOB coordinates: %s\n", cudaGetErrorString(e));
  }
  return 0;
}
