// Micro-benchmark: throughput of packed fp32 (FFMA2 / FADD2 / FMUL2) against scalar FFMA on sm_100a, alone and mixed with
// ALU-pipe work (FSEL / FMNMX), to decide whether k_gn should process two particles per thread in packed form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu && ./ffma2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

constexpr int CH = 8;  // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s0) {
  float x[2 * CH];
  u64 p[CH];
  const float m = 1.0f + s0 * 1e-7f, ad = s0 * 1e-9f;
#pragma unroll
  for (int i = 0; i < 2 * CH; i++) x[i] = s0 + (float)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < CH; i++) p[i] = pk(x[2 * i], x[2 * i + 1]);
  const u64 mm = pk(m, m), aa = pk(ad, ad);
  float sel = 0.f;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {  // scalar FFMA: 2*CH per trip
#pragma unroll
      for (int i = 0; i < 2 * CH; i++) x[i] = fmaf(x[i], m, ad);
    } else if (MODE == 1) {  // FFMA2: CH per trip (= 2*CH fmas)
#pragma unroll
      for (int i = 0; i < CH; i++) p[i] = fma2(p[i], mm, aa);
    } else if (MODE == 2) {  // scalar FFMA + as many FMNMX (ALU pipe)
#pragma unroll
      for (int i = 0; i < 2 * CH; i++) { x[i] = fmaf(x[i], m, ad); }
#pragma unroll
      for (int i = 0; i < 2 * CH; i++) { sel = fminf(sel + 0.0f, x[i]); }
    } else if (MODE == 3) {  // FFMA2 + FMNMX on both halves (same flops and same ALU work as mode 2)
#pragma unroll
      for (int i = 0; i < CH; i++) p[i] = fma2(p[i], mm, aa);
#pragma unroll
      for (int i = 0; i < CH; i++) { float a, b; upk(p[i], a, b); sel = fminf(sel, a); sel = fminf(sel, b); }
    } else if (MODE == 4) {  // FADD2 + FMUL2
#pragma unroll
      for (int i = 0; i < CH; i++) p[i] = (i & 1) ? add2(p[i], aa) : mul2(p[i], mm);
    } else if (MODE == 5) {  // FFMA2 with a broadcast operand built every trip ({s, s} from a scalar that changes)
      const float s = m + (float)it * 1e-9f;
      const u64 ss = pk(s, s);
#pragma unroll
      for (int i = 0; i < CH; i++) p[i] = fma2(p[i], ss, aa);
    }
  }
  float acc = sel;
  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int i = 0; i < 2 * CH; i++) acc += x[i];
  } else {
#pragma unroll
    for (int i = 0; i < CH; i++) { float a, b; upk(p[i], a, b); acc += a + b; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, int warps_per_sm) {
  int sm = 0, khz = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int threads = 256, blocks_per_sm = warps_per_sm * 32 / threads, iters = 20000;
  float *out;
  cudaMalloc(&out, (size_t)sm * blocks_per_sm * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sm * blocks_per_sm, threads>>>(out, 1000, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sm * blocks_per_sm, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)sm * warps_per_sm * 32 * (double)iters * 2 * CH;
  const double clk = ms * 1e-3 * khz * 1e3;
  printf("%-44s warps/SM %2d  %8.3f ms  %6.1f fp32-FMA lanes/clk/SM (at %d MHz nominal)  err=%s\n", name, warps_per_sm, ms, fmas / clk / sm, khz / 1000,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<0>("scalar FFMA", w);
    run<1>("FFMA2", w);
    run<2>("scalar FFMA + FMNMX (1:1)", w);
    run<3>("FFMA2 + FMNMX per half (same work)", w);
    run<4>("FADD2 / FMUL2", w);
    run<5>("FFMA2 with {s,s} built per trip", w);
  }
  return 0;
}
