// tma_probe.cu -- which tiled-TMA box shapes / coordinates are legal for FP64 without swizzle (debug aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, double* out, int c0, int c1, int bytes, int n) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = (s32(sm) + 1023u) & ~1023u;
  double* g = (double*)(sm + (base - s32(sm)));
  const uint32_t bar = base + 65536;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(base), "l"((uint64_t)&tm), "r"(bar), "r"(c0), "r"(c1), "r"(0) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = g[i];
}
int main(int argc, char** argv) {
  const int b0 = atoi(argv[1]), b1 = atoi(argv[2]), c0 = atoi(argv[3]), c1 = atoi(argv[4]);
  const int M = 1000, N = 100;
  std::vector<double> h(M * N); for (int i = 0; i < M * N; i++) h[i] = i;
  double *d, *o; CK(cudaMalloc(&d, h.size() * 8)); CK(cudaMalloc(&o, 65536)); CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  void* fn; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  typedef CUresult (*F)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm; cuuint64_t gd[3] = {(cuuint64_t)M, (cuuint64_t)N, 1}, gs[2] = {(cuuint64_t)M * 8, (cuuint64_t)M * N * 8}; cuuint32_t bx[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1}, es[3] = {1, 1, 1};
  CUresult r = ((F)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("box %dx%d: encode failed %d\n", b0, b1, (int)r); return 0; }
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000));
  const int n = b0 * b1;
  k<<<1, 128, 70000>>>(tm, o, c0, c1, n * 8, n);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("box %dx%d at (%d,%d): %s\n", b0, b1, c0, c1, cudaGetErrorString(e)); return 0; }
  std::vector<double> res(n); CK(cudaMemcpy(res.data(), o, n * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int j = 0; j < b1; j++) for (int i = 0; i < b0; i++) {
    const int gi = c0 + i, gj = c1 + j; const double want = (gi >= 0 && gi < M && gj >= 0 && gj < N) ? (double)(gi + gj * M) : 0.0;
    if (res[i + j * b0] != want) bad++;
  }
  printf("box %dx%d at (%d,%d): OK, %d mismatches\n", b0, b1, c0, c1, bad);
  return 0;
}
