// tc_probe.cu -- standalone probe for tcgen05.mma kind::tf32 descriptors (debug aid, not product code).
// One CTA: fills a 128x32 A tile and a 128x32 B tile in shared memory in the K-major SWIZZLE_128B canonical
// layout by plain stores, issues the 4 K=8 MMAs, reads the 128x128 FP32 accumulator back from TMEM and
// compares with the host result.  Variants are selected on the command line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { uint32_t idesc; uint32_t lbo, sbo, version, layout; int a_mn, b_mn; };

__global__ void probe(const float* A, const float* B, float* D, Cfg c) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  float* sA = (float*)gen;              // 16 KB
  float* sB = (float*)(gen + 16384);    // 16 KB
  const uint32_t bar = base + 32768, slot = base + 32768 + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  // element (row o, k) -> byte offset
  auto offK = [](int o, int k) { return o * 128 + ((((k >> 2) ^ (o & 7))) << 4) + ((k & 3) << 2); };          // K-major
  // MN-major for 32-bit types: SWIZZLE_128B_BASE32B -- 32-byte chunks XOR (k & 3), atoms of 4 k-rows
  auto offMN = [](int o, int k) { return (o >> 5) * 4096 + k * 128 + (((((o & 31) >> 3) ^ (k & 3))) << 5) + ((o & 7) << 2); };
  for (int i = tid; i < 128 * 32; i += blockDim.x) {
    const int o = i / 32, k = i % 32;
    *(float*)((uint8_t*)sA + (c.a_mn ? offMN(o, k) : offK(o, k))) = A[o * 32 + k];
    *(float*)((uint8_t*)sB + (c.b_mn ? offMN(o, k) : offK(o, k))) = B[o * 32 + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (tid == 0) {
    for (int ks = 0; ks < 4; ks++) {
      auto mk = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
        uint64_t d = 0;
        d |= (uint64_t)((addr >> 4) & 0x3FFF);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)c.version << 46;
        return d;
      };
      const uint64_t ad = c.a_mn ? (mk(base + ks * 1024, 4096, c.sbo) | ((uint64_t)c.layout << 61)) : (mk(base + ks * 32, 16, 1024) | ((uint64_t)2 << 61));
      const uint64_t bd = c.b_mn ? (mk(base + 16384 + ks * 1024, 4096, c.sbo) | ((uint64_t)c.layout << 61)) : (mk(base + 16384 + ks * 32, 16, 1024) | ((uint64_t)2 << 61));
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(c.idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  // everybody waits for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; j++) D[row * 128 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// TS mode: A operand lives in TMEM (lane = row m, column = k), written by tcgen05.st; B from shared memory.
__global__ void probe_ts(const float* A, const float* B, float* D, Cfg c) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  float* sB = (float*)(gen + 16384);
  const uint32_t bar = base + 32768, slot = base + 32768 + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto offK = [](int o, int k) { return o * 128 + ((((k >> 2) ^ (o & 7))) << 4) + ((k & 3) << 2); };
  auto offMN = [](int o, int k) { return (o >> 5) * 4096 + k * 128 + (((((o & 31) >> 3) ^ (k & 3))) << 5) + ((o & 7) << 2); };
  for (int i = tid; i < 128 * 32; i += blockDim.x) {
    const int o = i / 32, k = i % 32;
    *(float*)((uint8_t*)sB + (c.b_mn ? offMN(o, k) : offK(o, k))) = B[o * 32 + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t a_tmem = tmem + 128;   // A tile: columns 128..159
  if (warp < 4) {
    const int row = warp * 32 + lane;
    uint32_t r[32];
    for (int k = 0; k < 32; k++) r[k] = __float_as_uint(A[row * 32 + k]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(a_tmem + ((uint32_t)(warp * 32) << 16)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
          "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
          "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    for (int ks = 0; ks < 4; ks++) {
      auto mk = [&](uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
        uint64_t d = 0;
        d |= (uint64_t)((addr >> 4) & 0x3FFF);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)layout << 61;
        return d;
      };
      const uint64_t bd = c.b_mn ? mk(base + 16384 + ks * 1024, 4096, 512, 1) : mk(base + 16384 + ks * 32, 16, 1024, 2);
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(tmem), "r"(a_tmem + ks * c.lbo), "l"(bd), "r"(c.idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; j++) D[row * 128 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<float> A(128 * 32), B(128 * 32), D(128 * 128), W(128 * 128);
  for (int i = 0; i < 128 * 32; i++) { A[i] = (float)((i * 7) % 13) - 6; B[i] = (float)((i * 5) % 11) - 5; }
  for (int m = 0; m < 128; m++) for (int n = 0; n < 128; n++) {
    float s = 0; for (int k = 0; k < 32; k++) s += A[m * 32 + k] * B[n * 32 + k]; W[m * 128 + n] = s; }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
  auto idesc = [](int amn, int bmn, int afmt, int bfmt) {
    uint32_t d = 0; d |= 1u << 4; d |= (uint32_t)afmt << 7; d |= (uint32_t)bfmt << 10; d |= (uint32_t)amn << 15; d |= (uint32_t)bmn << 16;
    d |= (128u >> 3) << 17; d |= (128u >> 4) << 24; return d; };
  struct V { const char* name; Cfg c; };
  std::vector<V> vs = {
    {"K/K  ver1 sw128 lbo16 sbo1024", {idesc(0, 0, 2, 2), 16, 1024, 1, 2, 0, 0}},
    {"K/K  ver0 sw128", {idesc(0, 0, 2, 2), 16, 1024, 0, 2, 0, 0}},
    {"K/K  ver1 sw128 lbo0", {idesc(0, 0, 2, 2), 0, 1024, 1, 2, 0, 0}},
    {"MN/K  base32B sbo512", {idesc(1, 0, 2, 2), 16, 512, 1, 1, 1, 0}},
    {"K/MN  base32B sbo512", {idesc(0, 1, 2, 2), 16, 512, 1, 1, 0, 1}},
    {"MN/MN base32B sbo512", {idesc(1, 1, 2, 2), 16, 512, 1, 1, 1, 1}},
    {"MN/MN base32B sbo1024", {idesc(1, 1, 2, 2), 16, 1024, 1, 1, 1, 1}},
  };
  for (auto& v : vs) {
    CK(cudaMemset(dD, 0xFF, D.size() * 4));
    probe<<<1, 128, 40000>>>(dA, dB, dD, v.c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-36s : CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxd = 0; int zeros = 0;
    for (int i = 0; i < 128 * 128; i++) { maxd = fmax(maxd, fabs((double)D[i] - W[i])); zeros += (D[i] == 0.f); }
    printf("%-36s : max |diff| %.3f  zeros %d  D[0..3] = %.1f %.1f %.1f %.1f  want %.1f %.1f %.1f %.1f\n", v.name, maxd, zeros,
           D[0], D[1], D[2], D[3], W[0], W[1], W[2], W[3]);
  }
  // ---- TS mode (A in TMEM); cfg.lbo is reused as "TMEM columns per K=8 step" ----
  CK(cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
  std::vector<V> ts = {
    {"TS  A[tmem]/K   8 cols per k-step", {idesc(0, 0, 2, 2), 8, 0, 1, 2, 0, 0}},
    {"TS  A[tmem]/MN  8 cols per k-step", {idesc(0, 1, 2, 2), 8, 0, 1, 1, 0, 1}},
    {"TS  A[tmem]/K   4 cols per k-step", {idesc(0, 0, 2, 2), 4, 0, 1, 2, 0, 0}},
  };
  for (auto& v : ts) {
    CK(cudaMemset(dD, 0xFF, D.size() * 4));
    probe_ts<<<1, 128, 40000>>>(dA, dB, dD, v.c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-36s : CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxd = 0; int zeros = 0;
    for (int i = 0; i < 128 * 128; i++) { maxd = fmax(maxd, fabs((double)D[i] - W[i])); zeros += (D[i] == 0.f); }
    printf("%-36s : max |diff| %.3f  zeros %d  D[0..3] = %.1f %.1f %.1f %.1f  want %.1f %.1f %.1f %.1f\n", v.name, maxd, zeros,
           D[0], D[1], D[2], D[3], W[0], W[1], W[2], W[3]);
  }
  return 0;
}
