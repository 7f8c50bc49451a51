// popc_peak.cu -- measured POPC issue rate of one B200 SM (the ceiling of the xnor_popc engine).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/popc_peak tools/popc_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = seed * (threadIdx.x + 1 + i * 977); acc[i] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { acc[i] += __popc(x[i]); x[i] ^= acc[i]; }              // popc + add + xor  (3 instr, 1 popc)
      else if (MODE == 1) { acc[i] += __popc(~(x[i] ^ seed)); x[i] += 0x9e3779b9u; }  // xnor-popc-add + add (4 instr, 1 popc)
      else { acc[i] = (acc[i] ^ x[i]) + 0x9e3779b9u; x[i] = ~(x[i] ^ acc[i]); }      // no popc (LOP3/IADD only)
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r += acc[i] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  uint32_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  for (int mode = 0; mode < 3; mode++)
    for (int warps = 4; warps <= 32; warps *= 2) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      auto launch = [&]() { if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 12345u); else if (mode == 1) k<1><<<148, warps * 32>>>(d, iters, 12345u); else k<2><<<148, warps * 32>>>(d, iters, 12345u); };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(a); launch(); cudaEventRecord(b); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double ops = (double)iters * 8 * warps * 32;  // per SM
      int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
      printf("mode %d (%s) warps/SM %2d: %.3f ms -> %.1f lane-ops/clk/SM at %.0f MHz nominal (%.2f T lane-ops/s chip)\n", mode,
             mode == 0 ? "popc+add+xor" : mode == 1 ? "xnor+popc+add+add" : "lop3+iadd only", warps, ms, ops / (ms * 1e-3) / (clk * 1e3), clk / 1e3,
             ops * 148 / (ms * 1e-3) / 1e12);
    }
  return 0;
}
