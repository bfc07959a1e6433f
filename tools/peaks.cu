// peaks.cu -- register-resident issue-rate microbenchmarks for the B200 roofline denominators
// that MEASURED_PEAKS.json does not hold: FP64 DMMA (mma.sync f64 shapes),
// FP64 DFMA, FP32 FFMA, and the legacy mma.sync TF32 path.  Prints one JSON object.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/peaks tools/peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NACC = 8;     // independent accumulator tiles per warp (ILP)

// ---- DMMA m8n8k4: 256 FMA per warp instruction -------------------------------------------------
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
// ---- DMMA m16n8k4: 512 FMA -----------------------------------------------------------------------
__global__ void k_dmma1684(double* out, int iters, double a0, double b0) {
  double a[2] = {a0 + threadIdx.x * 1e-9, a0}, b = b0;
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 12345.678) out[0] = s;
}
// ---- DMMA m16n8k8: 1024 FMA ----------------------------------------------------------------------
__global__ void k_dmma1688(double* out, int iters, double a0, double b0) {
  double a[4] = {a0 + threadIdx.x * 1e-9, a0, a0, a0}, b[2] = {b0, b0};
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 12345.678) out[0] = s;
}
// ---- DMMA m16n8k16: 2048 FMA ---------------------------------------------------------------------
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
  double a[8], b[4];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = a0 + threadIdx.x * 1e-9 * j;
#pragma unroll
  for (int j = 0; j < 4; j++) b[j] = b0;
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                   "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 12345.678) out[0] = s;
}
// ---- DFMA / FFMA: 16 independent chains per thread ------------------------------------------------
template <typename T>
__global__ void k_fma(T* out, int iters, T a0, T b0) {
  T a = a0 + T(threadIdx.x) * T(1e-6), b = b0;
  T c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = T(i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = fma(a, c[i], b);
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  if (s == T(12345.678)) out[0] = s;
}
// ---- legacy mma.sync tf32 m16n8k8: 1024 FMA -------------------------------------------------------
__global__ void k_tf32mma(float* out, int iters, float a0, float b0) {
  unsigned a[4], b[2];
#pragma unroll
  for (int j = 0; j < 4; j++) a[j] = __float_as_uint(a0 + threadIdx.x * 1e-6f);
  b[0] = b[1] = __float_as_uint(b0);
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}
// sustained: loop back-to-back for ~secs seconds, return mean ms per launch
template <typename F>
static double sustained_ms(F launch, double secs, double est_ms) {
  int n = (int)(secs * 1000.0 / est_ms) + 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0)); for (int i = 0; i < n; i++) launch(); CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / n;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 1024));
  float* fout = (float*)dout;
  double secs = argc > 1 ? atof(argv[1]) : 2.0;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
  bool first = true;
  auto report = [&](const char* name, int warps, double fma_per_launch, double burst_ms, double sust_ms) {
    printf("%s {\"name\": \"%s\", \"warps_per_sm\": %d, \"tflops_burst\": %.3f, \"tflops_sustained\": %.3f, "
           "\"fma_per_clk_per_sm_at_1965MHz\": %.2f}", first ? "" : ",\n", name, warps,
           2 * fma_per_launch / burst_ms / 1e9, 2 * fma_per_launch / sust_ms / 1e9,
           fma_per_launch / (burst_ms * 1e-3) / sms / 1.965e9);
    first = false; fflush(stdout);
  };
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    int threads = 32 * (warps > 32 ? 32 : warps);
    int blocks = sms * (warps > 32 ? warps / 32 : 1);
    double nw = (double)blocks * threads / 32;
    { auto l = [&] { k_dmma884<<<blocks, threads>>>(dout, iters, 1.0, 1e-3); };
      double b = time_ms(l, 5); report("dmma_m8n8k4", warps, nw * iters * NACC * 256.0, b, sustained_ms(l, secs, b)); }
    { auto l = [&] { k_dmma1684<<<blocks, threads>>>(dout, iters / 2, 1.0, 1e-3); };
      double b = time_ms(l, 5); report("dmma_m16n8k4", warps, nw * (iters / 2) * NACC * 512.0, b, sustained_ms(l, secs, b)); }
    { auto l = [&] { k_dmma1688<<<blocks, threads>>>(dout, iters / 4, 1.0, 1e-3); };
      double b = time_ms(l, 5); report("dmma_m16n8k8", warps, nw * (iters / 4) * NACC * 1024.0, b, sustained_ms(l, secs, b)); }
    { auto l = [&] { k_dmma16816<<<blocks, threads>>>(dout, iters / 8, 1.0, 1e-3); };
      double b = time_ms(l, 5); report("dmma_m16n8k16", warps, nw * (iters / 8) * NACC * 2048.0, b, sustained_ms(l, secs, b)); }
    { auto l = [&] { k_fma<double><<<blocks, threads>>>(dout, iters, 0.999, 1e-3); };
      double b = time_ms(l, 5); report("dfma", warps, nw * 32.0 * iters * 16, b, sustained_ms(l, secs, b)); }
    { auto l = [&] { k_fma<float><<<blocks, threads>>>(fout, iters * 4, 0.999f, 1e-3f); };
      double b = time_ms(l, 5); report("ffma", warps, nw * 32.0 * iters * 4 * 16, b, sustained_ms(l, secs, b)); }
    { auto l = [&] { k_tf32mma<<<blocks, threads>>>(fout, iters, 1.0f, 1e-3f); };
      double b = time_ms(l, 5); report("mma_sync_tf32_m16n8k8", warps, nw * iters * NACC * 1024.0, b, sustained_ms(l, secs, b)); }
  }
  printf("\n]}\n");
  return 0;
}
