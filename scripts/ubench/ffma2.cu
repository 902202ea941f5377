// Micro-benchmark: issue rate of packed fp32 (FFMA2) vs scalar FFMA on sm_100a, and of 16-byte vs 32-byte gathers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu     Run on the GPU box.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float s) {
  float a[16]; uint64_t p[8];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = threadIdx.x * 0.001f + k;
#pragma unroll
  for (int k = 0; k < 8; k++) p[k] = pk(a[2 * k], a[2 * k + 1]);
  const uint64_t ss = pk(s, s), cc = pk(0.5f, 0.25f);
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 16; k++) a[k] = fmaf(a[k], s, 0.5f);
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) p[k] = fma2(p[k], ss, cc);
    }
  }
  float r = 0.f;
  if (MODE == 0) { for (int k = 0; k < 16; k++) r += a[k]; }
  else { for (int k = 0; k < 8; k++) { float x, y; upk(p[k], x, y); r += x + y; } }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// gathers: each lane reads REC bytes at a pseudo-random index within a window (emulates neighbour gathers)
template <int REC>
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ src, const uint32_t* __restrict__ idx, float* out, int per_thread, uint32_t n) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int k = 0; k < per_thread; k++) {
    const uint32_t j = idx[(size_t)k * n + t];
    if (REC == 16) { const float4 v = __ldg(&src[j]); acc += v.x + v.w; }
    else {
      float4 v0, v1;
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w), "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "l"(src + 2 * (size_t)j));
      acc += v0.x + v1.w;
    }
  }
  out[t] = acc;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 64 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 8;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k_fma<0><<<grid, 256>>>(out, iters, 0.999f); else k_fma<1><<<grid, 256>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fmas = (double)grid * 256 * iters * 16;
      if (rep) printf("%s: %.3f ms  %.2f TFMA/s (lane-FMAs)  %.2f Tinstr-lanes/s\n", mode ? "FFMA2" : "FFMA ", ms, fmas / ms * 1e-9, fmas / (mode ? 2 : 1) / ms * 1e-9);
    }
  }
  // gather test: n particles, each thread gathers 64 neighbours at index t + small pseudo-random offsets (27-cell-like window)
  const uint32_t n = 1u << 22; const int per = 64;
  std::vector<uint32_t> hidx((size_t)per * n);
  uint32_t s = 12345u;
  for (int k = 0; k < per; k++) for (uint32_t t = 0; t < n; t++) {
    s = s * 1664525u + 1013904223u;
    const int run = (int)((s >> 8) % 9) - 4;                 // one of 9 z-runs, ~1800 particles apart
    const int off = (int)((s >> 16) % 81) - 40;
    long j = (long)t + run * 1800L + off;
    if (j < 0) j = 0; if (j >= n) j = n - 1;
    hidx[(size_t)k * n + t] = (uint32_t)j;
  }
  uint32_t* didx; float4* src; cudaMalloc(&didx, hidx.size() * 4); cudaMalloc(&src, (size_t)n * 32);
  cudaMemcpy(didx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice); cudaMemset(src, 0, (size_t)n * 32);
  float* o2; cudaMalloc(&o2, n * 4);
  for (int rec = 16; rec <= 32; rec += 16)
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (rec == 16) k_gather<16><<<n / 256, 256>>>(src, didx, o2, per, n); else k_gather<32><<<n / 256, 256>>>(src, didx, o2, per, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("gather %2d B: %.3f ms  %.2f Ggathers/s  (%s)\n", rec, ms, (double)n * per / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
