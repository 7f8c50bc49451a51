// tma_probe.cu -- which TMA tile-load variants of a 4-byte-pixel image work (rank, dtype, swizzle).  nvcc -arch=sm_100a -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap m, uint32_t* out, int nwords, int c0, int c1) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(nwords * 4));
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(sm)),
                   "l"((uint64_t)&m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(0) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(s32(sm)),
                   "l"((uint64_t)&m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar)));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(sm)[i];
}
int main(int argc, char** argv) {
  const int C0 = argc > 1 ? atoi(argv[1]) : -2, only = argc > 2 ? atoi(argv[2]) : -1;
  const int X = 12, Y = 8, N = 2, BW = 16, BH = 11;
  std::vector<uint32_t> h(X * Y * N);
  for (size_t i = 0; i < h.size(); i++) h[i] = 1000 + i;
  uint32_t *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 8192);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 4; variant++) {
    if (only >= 0 && variant != only) continue;
    const int rank = (variant & 1) ? 4 : 3;
    const bool u8 = variant & 2;
    CUtensorMap m;
    cuuint64_t gd[4] = {(cuuint64_t)(u8 ? X * 4 : X), Y, N, 1}, gs[3] = {X * 4, X * Y * 4, (cuuint64_t)X * Y * N * 4};
    cuuint32_t bx[4] = {(cuuint32_t)(u8 ? BW * 4 : BW), BH, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, d, gd, gs, bx, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d rank %d %s: encode=%d ", variant, rank, u8 ? "u8" : "u32", (int)r);
    if (r) { printf("\n"); continue; }
    cudaMemset(o, 0xFF, 8192);
    if (rank == 3) k<3><<<1, 128, 16384>>>(m, o, BW * BH, u8 ? 4 * C0 : C0, -2);
    else k<4><<<1, 128, 16384>>>(m, o, BW * BH, u8 ? 4 * C0 : C0, -2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e) { printf("\n"); return 1; }
    std::vector<uint32_t> g(BW * BH);
    cudaMemcpy(g.data(), o, BW * BH * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < BH; y++)
      for (int x = 0; x < BW; x++) {
        const int sx = x + C0, sy = y - 2;
        const uint32_t want = (sx >= 0 && sx < X && sy >= 0 && sy < Y) ? 1000 + sy * X + sx : 0;
        bad += g[y * BW + x] != want;
      }
    printf("mismatches=%d\n", bad);
  }
  return 0;
}
