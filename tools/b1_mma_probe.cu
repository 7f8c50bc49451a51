#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a0 = seed ^ threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  int c[4][4] = {};
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  uint32_t s = 0;
  for (int j = 0; j < 4; j++) for (int q = 0; q < 4; q++) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// layout check: A row r = all ones in bits [0, r), B col n = all zeros -> xor popc = r  => C[r][n] = r
__global__ void lay(int* out) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  // A[r][kbit]: bit set iff kbit < 8*r + 3 ; B[k][n] bit set iff kbit < n  (xor popc = |8r+3 - n|)
  auto abits = [&](int r, int k0) { uint32_t v = 0; for (int b = 0; b < 32; b++) if (k0 + b < 8 * r + 3) v |= 1u << b; return v; };
  auto bbits = [&](int n, int k0) { uint32_t v = 0; for (int b = 0; b < 32; b++) if (k0 + b < n) v |= 1u << b; return v; };
  uint32_t a0 = abits(g, t * 32), a1 = abits(g + 8, t * 32), a2 = abits(g, 128 + t * 32), a3 = abits(g + 8, 128 + t * 32);
  uint32_t b0 = bbits(g, t * 32), b1 = bbits(g, 128 + t * 32);
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  out[(g) * 8 + 2 * t] = c0; out[(g) * 8 + 2 * t + 1] = c1; out[(g + 8) * 8 + 2 * t] = c2; out[(g + 8) * 8 + 2 * t + 1] = c3;
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  int* dl; cudaMalloc(&dl, 128 * 4);
  lay<<<1, 32>>>(dl);
  int h[128]; cudaMemcpy(h, dl, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < 16; r++) for (int n = 0; n < 8; n++) { int want = abs(8 * r + 3 - n); if (h[r * 8 + n] != want) { if (bad < 8) printf("layout mismatch r=%d n=%d got %d want %d\n", r, n, h[r * 8 + n], want); bad++; } }
  printf("layout check: %d mismatches (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
  for (int warps : {4, 8, 16}) {
    const int iters = 20000;
    k<<<148, warps * 32>>>(d, 100, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148, warps * 32>>>(d, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = 148.0 * warps * iters * 4;
    printf("%2d warps/SM: %.3f ms, %.1f G mma/s, %.2f Pbit-MAC/s, %.1f clk per mma per SM (1.9 GHz), err=%s\n", warps, ms, mmas / ms / 1e6,
           mmas * 16 * 8 * 256 / ms / 1e12, ms * 1e-3 * 1.9e9 / (warps * iters * 4.0), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
