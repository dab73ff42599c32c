// Microbenchmark: cost model of fire-and-forget fp32 reductions (REDG) on B200.
// Each warp issues NITER reduction instructions to pseudo-random rows of a table.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void red4(float* p, float a) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red2(float* p, float a) {
  asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red1(float* p, float a) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
// mode 0: v4, lane pairs share a 32B row (16 sectors / instr)
// mode 1: scalar, 16 active lanes, 16 sectors
// mode 2: scalar, 32 lanes, 32 sectors
// mode 3: v4, 32 lanes, 32 distinct sectors (16B each)
// mode 4: v4, 8 lanes cover one 128B line (4 lines / instr)
// mode 5: v4, 4 lanes cover 64B (8 half-lines / instr)
// mode 6: v2, 4 lanes share a 32B row (8 sectors)
// mode 7: loads instead (ld.cg v4 paired, 16 sectors) for comparison
template <int MODE>
__global__ void k(float* tab, uint32_t rows32 /*number of 32B rows*/, int niter, float* sink) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  for (int i = 0; i < niter; i++) {
    const uint32_t base = hash(gw * 7919u + i * 104729u);
    if (MODE == 0) { uint32_t r = hash(base + (lane >> 1)) % rows32; red4(tab + (size_t)r * 8 + (lane & 1) * 4, 1e-9f); }
    if (MODE == 1) { if (lane < 16) { uint32_t r = hash(base + lane) % rows32; red1(tab + (size_t)r * 8, 1e-9f); } }
    if (MODE == 2) { uint32_t r = hash(base + lane) % rows32; red1(tab + (size_t)r * 8, 1e-9f); }
    if (MODE == 3) { uint32_t r = hash(base + lane) % rows32; red4(tab + (size_t)r * 8, 1e-9f); }
    if (MODE == 4) { uint32_t r = hash(base + (lane >> 3)) % (rows32 / 4); red4(tab + (size_t)r * 32 + (lane & 7) * 4, 1e-9f); }
    if (MODE == 5) { uint32_t r = hash(base + (lane >> 2)) % (rows32 / 2); red4(tab + (size_t)r * 16 + (lane & 3) * 4, 1e-9f); }
    if (MODE == 6) { uint32_t r = hash(base + (lane >> 2)) % rows32; red2(tab + (size_t)r * 8 + (lane & 3) * 2, 1e-9f); }
    if (MODE == 8) { uint32_t r = hash(base + lane) % rows32; red1(tab + r, 1e-9f); }
    if (MODE == 9) { uint32_t r = hash(base + lane) % rows32; acc += __ldcg(tab + r); }
    if (MODE == 10) { uint32_t r = hash(base + lane) % rows32; acc += __ldcg(tab + r); red1(tab + r, 1e-9f); }
    if (MODE == 11) { uint32_t r = hash(base + lane) % rows32; acc += __ldcg(tab + (size_t)r * 8); red1(tab + (size_t)r * 8, 1e-9f); }
    if (MODE == 7) { uint32_t r = hash(base + (lane >> 1)) % rows32; float4 v = __ldcg(reinterpret_cast<const float4*>(tab + (size_t)r * 8 + (lane & 1) * 4)); acc += v.x + v.w; }
  }
  if (acc == 123.456f) *sink = acc;
}
template <int MODE>
void run(const char* name, float* tab, uint32_t rows32, int grid, int block, int niter, float* sink) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<grid, block>>>(tab, rows32, niter, sink);
  cudaEventRecord(a);
  k<MODE><<<grid, block>>>(tab, rows32, niter, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double instr = (double)grid * (block / 32) * niter;
  int sms = grid < 148 ? grid : 148;
  printf("%-44s rows=%8u grid=%4d: %8.1f us  %7.2f G instr/s  %6.1f cyc/instr/SM(@1.9GHz)\n", name, rows32, grid, ms * 1e3,
         instr / ms / 1e6, ms * 1e-3 * 1.9e9 * sms / instr);
}
int main() {
  float *tab, *sink; size_t bytes = 512ull << 20;
  cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes); cudaMalloc(&sink, 4);
  const int niter = 256, block = 256;
  for (uint32_t rows32 : {9746u, 82248u}) {
    int grid = 148 * 4;
    run<2>("scalar RED 32 lanes, 32B stride", tab, rows32, grid, block, niter, sink);
    run<8>("scalar RED 32 lanes, contiguous 4B", tab, rows32, grid, block, niter, sink);
    run<9>("scalar LD  32 lanes, contiguous 4B", tab, rows32, grid, block, niter, sink);
    run<10>("LD+RED same word, contiguous 4B", tab, rows32, grid, block, niter, sink);
    run<11>("LD+RED same word, 32B stride", tab, rows32, grid, block, niter, sink);
    run<0>("v4 paired RED (16 sectors)", tab, rows32, grid, block, niter, sink);
    run<7>("v4 paired LD (16 sectors)", tab, rows32, grid, block, niter, sink);
  }
  return 0;
}
