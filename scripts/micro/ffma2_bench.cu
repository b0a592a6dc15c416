// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
constexpr int ITERS = 4096, NACC = 8;
__global__ void k_scalar(float *out, float a, float b) {
  float acc[NACC];
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3f + i;
  float x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma1(acc[i], x, y);
  float s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float *out, float a, float b) {
  uint64_t acc[NACC];
  for (int i = 0; i < NACC; i++) acc[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i);
  uint64_t x = pk(a + threadIdx.x, a), y = pk(b, b);
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma2(acc[i], x, y);
  float s = 0; for (int i = 0; i < NACC; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    double inst = 148.0 * 8 * 256 * ITERS * NACC;
    printf("scalar FFMA : %.3f ms  %.1f G lane-inst/s  %.1f TFLOP/s\n", ms, inst / ms / 1e6, 2 * inst / ms / 1e9);
    cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2: %.3f ms  %.1f G lane-inst/s  %.1f TFLOP/s\n", ms, inst / ms / 1e6, 4 * inst / ms / 1e9);
  }
  return 0;
}
// ---- mixed: NP packed + NS scalar independent accumulators per loop body ----
template <int NP, int NS>
__global__ void k_mixed(float *out, float a, float b) {
  uint64_t accp[NP > 0 ? NP : 1];
  float accs[NS > 0 ? NS : 1];
  for (int i = 0; i < NP; i++) accp[i] = pk(threadIdx.x * 1e-3f + i, 1.f + i);
  for (int i = 0; i < NS; i++) accs[i] = threadIdx.x * 1e-3f + i;
  uint64_t x2 = pk(a + threadIdx.x, a), y2 = pk(b, b);
  float x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < (NP > NS ? NP : NS); i++) {
      if (i < NP) accp[i] = fma2(accp[i], x2, y2);
      if (i < NS) accs[i] = fma1(accs[i], x, y);
    }
  }
  float s = 0;
  for (int i = 0; i < NP; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(accp[i])); s += lo + hi; }
  for (int i = 0; i < NS; i++) s += accs[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NP, int NS>
void run_mixed(float *out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  k_mixed<NP, NS><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e0); k_mixed<NP, NS><<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  double results = 148.0 * 8 * 256 * ITERS * (2.0 * NP + NS);
  double insts = 148.0 * 8 * 256 * ITERS * (NP + NS);
  printf("mixed %d packed + %d scalar: %.3f ms  %.1f G results/s  %.1f G lane-inst/s\n", NP, NS, ms, results / ms / 1e6, insts / ms / 1e6);
}
struct Init { Init() { float *out; cudaMalloc(&out, 148 * 8 * 256 * 4); run_mixed<8,0>(out); run_mixed<0,8>(out); run_mixed<4,4>(out); run_mixed<4,8>(out); run_mixed<2,8>(out); run_mixed<6,4>(out); } } g_init;
