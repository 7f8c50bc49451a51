// tools/umma_peak.cu -- measures the dense INT8 tensor-core ceiling of a B200 with bare tcgen05.mma kind::i8
// loops (no global traffic): the denominator SURVEY.md 8(d) asks for next to the 4.5 POPS spec figure, and
// the answer to "which instruction shape can the shared-memory operand path feed?".
//   variant A: cta_group::1, M=128, N in {64,128,256}   (A and B operands from shared memory, K-major SW128)
//   variant B: cta_group::2, M=256 (128 per CTA), N in {128,256}, B split across the CTA pair
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/umma_peak tools/umma_peak.cu
#include <cstdio>
#include <cstdlib>

#include "../simple_image_compression_network_b200/csrc/fcb_sm100.cuh"
using namespace fcb::sm100;

#define CK(x)                                                                                                          \
  do {                                                                                                                 \
    cudaError_t e = (x);                                                                                               \
    if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } \
  } while (0)

__device__ __forceinline__ void fill_smem(uint8_t* p, int bytes, uint32_t seed) {
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) {
    uint32_t x = (i + 1) * 2654435761u ^ seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    reinterpret_cast<uint32_t*>(p)[i] = x;
  }
}

__global__ void __launch_bounds__(128, 1) peak1(int n, int iters, int shift_rows, int pattern) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 65536;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  fill_smem(sA, 65536 + 32768, blockIdx.x);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_i8(128, n, 0, 1);
    const uint64_t adesc = make_smem_desc(smem_u32(sA) + shift_rows * 128, 128), bdesc = make_smem_desc(smem_u32(sB), 128);
    const int batches = iters / 16;
    if (pattern) mbar_arrive(&bar[2]);  // a completed barrier to poll, like the full[] barriers of the real kernel
    for (int b = 0; b < batches; b++) {
      for (int i = 0; i < 16; i++) {
        if (pattern == 1) {  // the conv kernel's K-block: poll a barrier, 2 M-blocks x 4 k-steps, commit
          mbar_wait(&bar[2], 0);
          tc_fence_after();
          const uint64_t a2 = make_smem_desc(smem_u32(sA) + (shift_rows + (i % 3) * 50 + (i % 5)) * 128, 128);
#pragma unroll
          for (int k = 0; k < 4; k++) umma_i8(tmem, a2 + 2 * k, bdesc + 2 * k, idesc, 1u);
          const uint64_t a3 = a2 + 1024;
#pragma unroll
          for (int k = 0; k < 4; k++) umma_i8(tmem + 128u, a3 + 2 * k, bdesc + 2 * k, idesc, 1u);
          i++;
          continue;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) umma_i8(tmem + (uint32_t)((i & 1) * 256), adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
      }
      umma_commit(&bar[b & 1]);
      if (b > 0) mbar_wait(&bar[(b - 1) & 1], ((b - 1) >> 1) & 1);
    }
    mbar_wait(&bar[(batches - 1) & 1], ((batches - 1) >> 1) & 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}


// ---- issue-pattern study: warp-uniform loop, one elected lane issues (the structure of the conv kernel) -------------
// mode bits: 1 = A start shifted by one 128-byte row, 2 = commit after every 8 MMAs, 4 = poll a completed mbarrier +
// tcgen05.fence before every 8 MMAs, 8 = poll two barriers, 16 = A/B addresses vary per K-block like the conv kernel,
// 32 = __syncwarp after each K-block
__global__ void __launch_bounds__(128, 1) peak3(int iters, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 65536;  // 3 x 16 KB "weight stages"
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  fill_smem(sA, 65536 + 49152, blockIdx.x);
  if (threadIdx.x == 0) { for (int i = 0; i < 6; i++) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
    const uint32_t idesc = make_idesc_i8(128, 128, 0, 1);
    const uint32_t a0 = smem_u32(sA) + ((mode & 1) ? 128 : 0), b0 = smem_u32(sB);
    if (threadIdx.x == 0) { mbar_arrive(&bar[2]); mbar_arrive(&bar[3]); }
    __syncwarp();
    const int nkb = iters / 2;  // K-blocks of 8 MMAs
    for (int kb = 0; kb < nkb; kb++) {
      if (mode & 4) { mbar_wait(&bar[2], 0); tc_fence_after(); }
      if (mode & 8) { mbar_wait(&bar[3], 0); tc_fence_after(); }
      uint32_t aa = a0, bb = b0;
      if (mode & 16) { aa += ((kb % 9) * 50 + (kb % 3)) * 128; bb += (kb % 3) * 16384; }
      const uint64_t adesc = make_smem_desc(aa, 128), bdesc = make_smem_desc(bb, 128);
      if (elect_one_sync()) {
#pragma unroll
        for (int mb = 0; mb < 2; mb++) {
          const uint64_t ad = adesc + (uint64_t)(mb * 1024);
          const uint32_t dt = tmem + (uint32_t)(mb * 128 + (kb & 1) * 256);
          umma_i8(dt, ad, bdesc, idesc, 1u);
          umma_i8(dt, ad + 2, bdesc + 2, idesc, 1u);
          umma_i8(dt, ad + 4, bdesc + 4, idesc, 1u);
          umma_i8(dt, ad + 6, bdesc + 6, idesc, 1u);
        }
        if (mode & 2) umma_commit(&bar[4]);
        if ((kb & 7) == 7) umma_commit(&bar[(kb >> 3) & 1]);
      }
      if (mode & 32) __syncwarp();
      if ((kb & 7) == 7 && kb >= 15) mbar_wait(&bar[((kb >> 3) - 1) & 1], (((kb >> 3) - 1) >> 1) & 1);
    }
    const int last = (nkb >> 3) - 1;
    mbar_wait(&bar[last & 1], (last >> 1) & 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*slot, 512);
}


// ---- operand-alignment study: M=128 x N=256, B (N-side) and/or A start shifted by whole 128-byte rows ----------------
__global__ void __launch_bounds__(128, 1) peak4(int iters, int a_shift, int b_shift, int n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;           // 128 rows (+ slack)
  uint8_t* sB = smem + 32768;   // 256 rows (+ slack)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  fill_smem(sA, 32768 + 65536, blockIdx.x);
  if (threadIdx.x == 0) { for (int i = 0; i < 2; i++) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
    const uint32_t idesc = make_idesc_i8(128, n, 1, 0);
    const uint64_t adesc = make_smem_desc(smem_u32(sA) + a_shift * 128, 128), bdesc = make_smem_desc(smem_u32(sB) + b_shift * 128, 128);
    const int nkb = iters;  // K-blocks of 4 MMAs
    for (int kb = 0; kb < nkb; kb++) {
      if (elect_one_sync()) {
        const uint32_t dt = tmem + (uint32_t)((kb & 1) * 256);
        umma_i8(dt, adesc, bdesc, idesc, 1u);
        umma_i8(dt, adesc + 2, bdesc + 2, idesc, 1u);
        umma_i8(dt, adesc + 4, bdesc + 4, idesc, 1u);
        umma_i8(dt, adesc + 6, bdesc + 6, idesc, 1u);
        if ((kb & 15) == 15) umma_commit(&bar[(kb >> 4) & 1]);
      }
      __syncwarp();
      if ((kb & 15) == 15 && kb >= 31) mbar_wait(&bar[((kb >> 4) - 1) & 1], (((kb >> 4) - 1) >> 1) & 1);
    }
    const int last = (nkb >> 4) - 1;
    mbar_wait(&bar[last & 1], (last >> 1) & 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*slot, 512);
}

// ---- cta_group::2 -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) peak2(int n, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;           // this CTA's 128 rows of A
  uint8_t* sB = smem + 16384;   // this CTA's n/2 rows of B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const uint32_t rank = cluster_ctarank();
  fill_smem(sA, 16384 + 32768, blockIdx.x);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_i8(256, n, 0, 1);
    const uint64_t adesc = make_smem_desc(smem_u32(sA), 128), bdesc = make_smem_desc(smem_u32(sB), 128);
    const int batches = iters / 16;
    for (int b = 0; b < batches; b++) {
      for (int i = 0; i < 16; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)((i & 1) * 256)),
              "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u)
              : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&bar[b & 1])),
                   "h"((uint16_t)1)
                   : "memory");
      if (b > 0) mbar_wait(&bar[(b - 1) & 1], ((b - 1) >> 1) & 1);
    }
    mbar_wait(&bar[(batches - 1) & 1], ((batches - 1) >> 1) & 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int iters = argc > 1 ? atoi(argv[1]) : 16384;
  const int smem = 65536 + 32768 + 1024 + 64;
  CK(cudaFuncSetAttribute(peak1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(peak2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("device %s, %d SMs, iters %d (x4 k-steps of K=32)\n", prop.name, sms, iters);
  const int ns[3] = {64, 128, 256};
  for (int rep = 0; rep < 2; rep++)
    for (int n : ns) {
      peak1<<<sms, 128, smem>>>(n, 64, 0, 0);  // warm
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      peak1<<<sms, 128, smem>>>(n, iters, 0, 0);
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("peak1 n=%d failed: %s\n", n, cudaGetErrorString(e)); return 3; }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = 2.0 * sms * (double)iters * 4 * 128.0 * n * 32.0;
      printf("cta_group::1 M=128 N=%3d : %8.3f ms  %8.1f TOP/s  (%.1f clk/MMA at 1.965 GHz)\n", n, ms, ops / ms / 1e9,
             ms * 1e-3 * 1.965e9 / (iters * 4.0));
    }
  {
    const int shifts[6] = {0, 1, 2, 4, 8, 51};
    for (int pat = 0; pat < 2; pat++)
      for (int sh : shifts) {
        peak1<<<sms, 128, smem>>>(128, 64, sh, pat);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        peak1<<<sms, 128, smem>>>(128, iters, sh, pat);
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift test failed: %s\n", cudaGetErrorString(e)); return 3; }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = 2.0 * sms * (double)iters * 4 * 128.0 * 128 * 32.0;
        printf("N=128 A start shifted by %2d rows, pattern %d : %8.3f ms  %8.1f TOP/s  (%.1f clk/MMA at 1.965 GHz)\n", sh, pat, ms,
               ops / ms / 1e9, ms * 1e-3 * 1.965e9 / (iters * 4.0));
      }
  }
  {
    const int smem3 = 65536 + 49152 + 1024 + 128;
    CK(cudaFuncSetAttribute(peak3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    const int modes[] = {0, 1, 2, 4, 6, 12, 14, 16, 17, 30, 62, 63};
    for (int mode : modes) {
      peak3<<<sms, 128, smem3>>>(128, mode);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      peak3<<<sms, 128, smem3>>>(iters, mode);
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("peak3 mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 3; }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = 2.0 * sms * (double)iters * 4 * 128.0 * 128 * 32.0;
      printf("uniform-issue M=128 N=128 mode %2d : %8.3f ms  %8.1f TOP/s  (%.1f clk/MMA at 1.965 GHz)\n", mode, ms, ops / ms / 1e9,
             ms * 1e-3 * 1.965e9 / (iters * 4.0));
    }
  }
  {
    const int smem4 = 32768 + 65536 + 1024 + 128;
    CK(cudaFuncSetAttribute(peak4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
    const int cfg[][3] = {{0, 0, 256}, {0, 1, 256}, {0, 2, 256}, {0, 4, 256}, {0, 8, 256}, {0, 51, 256}, {1, 0, 256}, {3, 5, 256},
                          {0, 0, 128}, {0, 1, 128}, {1, 0, 128}, {0, 0, 64}, {0, 3, 64}};
    for (auto& c : cfg) {
      peak4<<<sms, 128, smem4>>>(64, c[0], c[1], c[2]);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      peak4<<<sms, 128, smem4>>>(iters, c[0], c[1], c[2]);
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("peak4 failed: %s\n", cudaGetErrorString(e)); return 3; }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = 2.0 * sms * (double)iters * 4 * 128.0 * c[2] * 32.0;
      printf("align study M=128 N=%3d A-shift %2d rows B-shift %2d rows : %8.3f ms  %8.1f TOP/s  (%.1f clk/MMA at 1.965 GHz)\n", c[2], c[0],
             c[1], ms, ops / ms / 1e9, ms * 1e-3 * 1.965e9 / (iters * 4.0));
    }
  }
  const int ns2[2] = {128, 256};
  for (int n : ns2) {
    peak2<<<sms, 128, smem>>>(n, 64);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("peak2 n=%d failed: %s\n", n, cudaGetErrorString(e)); return 3; }
    CK(cudaEventRecord(e0));
    peak2<<<sms, 128, smem>>>(n, iters);
    CK(cudaEventRecord(e1));
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("peak2 n=%d failed: %s\n", n, cudaGetErrorString(e)); return 3; }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double ops = 2.0 * (sms / 2) * (double)iters * 4 * 256.0 * n * 32.0;
    printf("cta_group::2 M=256 N=%3d : %8.3f ms  %8.1f TOP/s  (%.1f clk/MMA at 1.965 GHz)\n", n, ms, ops / ms / 1e9,
           ms * 1e-3 * 1.965e9 / (iters * 4.0));
  }
  return 0;
}
