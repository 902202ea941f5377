// Micro-benchmark: what decides the cost of a divergent 16-byte (and 8-byte) gather in the L1 data stage of sm_100a?
// Every warp gathers records from a window around its own base; the index patterns differ only in how the 32 lanes'
// records are spread over 128-byte lines and over the 16-byte "chunk offsets" inside a line ((j mod 8) for float4 records).
//   0  coalesced            j = lane                               (4 lines, every chunk offset 4 times)
//   1  random               j uniform in the window
//   2  random lines, chunk offset = lane mod 8     (every quarter-warp covers all 8 offsets)
//   3  random lines, chunk offset = lane / 4       (every quarter-warp sits on 2 offsets; the warp as a whole is balanced)
//   4  random lines, chunk offset = 0 for all lanes
//   5  random inside 4 lines (32 records)
//   6  random inside 16 lines (128 records)
//   7  quarter q of the warp reads line q, random offsets inside it (4 lines per gather, like 0, but conflicts in a quarter)
//   8  8 distinct lines per quarter (32 lines), chunk offset = lane mod 8
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_banks gather_banks.cu
// Run:   ./gather_banks            (times)   |   ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,
//        l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum ./gather_banks 1     (one launch per pattern)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static const int WINDOW = 512;      // records per warp window (32 KB of float4 per 4 warps ... stays in L1/L2)
static const int ROWS = 64;         // gathers per lane per pass over the pattern table
static const int NPAT = 9;

template <int REC>                  // REC = 16: float4 records, 8: float2 records (same indices, half the bytes)
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ src, const uint16_t* __restrict__ pat, float* out, int reps, uint32_t n_rec) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = t >> 5;
  const uint32_t base = (uint32_t)(((uint64_t)warp * 1237u * 8u) % (n_rec - WINDOW)) & ~7u;      // line-aligned window
  float acc = 0.f;
  for (int r = 0; r < reps; r++) {
#pragma unroll 4
    for (int k = 0; k < ROWS; k++) {
      const uint32_t j = base + pat[k * 32 + lane];
      if (REC == 16) { const float4 v = __ldg(src + j); acc += v.x + v.w; }
      else { const float2 v = __ldg(reinterpret_cast<const float2*>(src) + j); acc += v.x + v.y; }
    }
  }
  out[t] = acc;
}

static uint32_t rnd(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }

int main(int argc, char** argv) {
  const int once = argc > 1 ? atoi(argv[1]) : 0;
  const uint32_t n_rec = 1u << 22;                  // 64 MB of float4
  float4* src; float* out; uint16_t* pat;
  cudaMalloc(&src, sizeof(float4) * n_rec); cudaMemset(src, 0, sizeof(float4) * n_rec);
  const int blocks = 148 * 6 * 4, threads = 256;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&pat, sizeof(uint16_t) * ROWS * 32);
  const char* names[NPAT] = {"coalesced", "random in 512 records", "random lines, offset = lane%8", "random lines, offset = lane/4", "random lines, offset 0",
                             "random in 4 lines", "random in 16 lines", "quarter q -> line q, random offsets", "32 distinct lines, offset = lane%8"};
  uint64_t seed = 12345;
  for (int rec = 16; rec >= 8; rec -= 8)
    for (int p = 0; p < NPAT; p++) {
      std::vector<uint16_t> h(ROWS * 32);
      for (int k = 0; k < ROWS; k++)
        for (int l = 0; l < 32; l++) {
          const uint32_t nl = WINDOW / 8;           // lines in the window
          uint32_t j = 0;
          switch (p) {
            case 0: j = l + 32 * (rnd(seed) % 1); break;
            case 1: j = rnd(seed) % WINDOW; break;
            case 2: j = (rnd(seed) % nl) * 8 + (l & 7); break;
            case 3: j = (rnd(seed) % nl) * 8 + (l >> 2); break;
            case 4: j = (rnd(seed) % nl) * 8; break;
            case 5: j = rnd(seed) % 32; break;
            case 6: j = rnd(seed) % 128; break;
            case 7: j = (l >> 3) * 8 + rnd(seed) % 8; break;
            case 8: j = (uint32_t)l * 8 + (l & 7); break;
          }
          h[k * 32 + l] = (uint16_t)j;
        }
      cudaMemcpy(pat, h.data(), sizeof(uint16_t) * ROWS * 32, cudaMemcpyHostToDevice);
      const int reps = once ? 2 : 16;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      if (!once) { if (rec == 16) k_gather<16><<<blocks, threads>>>(src, pat, out, reps, n_rec); else k_gather<8><<<blocks, threads>>>(src, pat, out, reps, n_rec); }
      cudaEventRecord(e0);
      if (rec == 16) k_gather<16><<<blocks, threads>>>(src, pat, out, reps, n_rec); else k_gather<8><<<blocks, threads>>>(src, pat, out, reps, n_rec);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double gathers = (double)blocks * threads / 32 * ROWS * reps;
      printf("rec %2d B  pattern %d  %-38s %8.3f ms  %7.2f cycles/SM per warp-gather (at 1.965 GHz)  err=%s\n", rec, p, names[p], ms,
             ms * 1e-3 * 1.965e9 * 148 / gathers, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
