// fcb_synth.cu -- counter-based synthetic bytes on the device: byte i = splitmix64(seed ^ (offset+i)) & mask.
// Bit-identical to simple_image_compression_network_b200/synth.py so host- and device-generated tensors agree
// (SURVEY.md 8(d)); used to fill HBM-resident benchmark inputs without a 50 GB host copy.
#include "fcb_internal.h"

namespace fcb {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void synth_fill_kernel(uint8_t* __restrict__ dst, size_t n, uint64_t seed, uint32_t mask, uint64_t offset) {
  // 16 bytes per thread per step, coalesced 128-bit stores
  const size_t n16 = n / 16;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n16; v += (size_t)gridDim.x * blockDim.x) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t acc = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) acc |= ((uint32_t)(splitmix64(seed ^ (offset + v * 16 + j * 4 + b)) & mask & 0xFFu)) << (8 * b);
      w[j] = acc;
    }
    reinterpret_cast<uint4*>(dst)[v] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (blockIdx.x == 0)
    for (size_t i = n16 * 16 + threadIdx.x; i < n; i += blockDim.x) dst[i] = (uint8_t)(splitmix64(seed ^ (offset + i)) & mask & 0xFFu);
}

int synth_fill(void* d_ptr, size_t n_bytes, uint64_t seed, uint32_t mask, uint64_t offset, cudaStream_t st) {
  if (!n_bytes) return FCB_OK;
  if (((uintptr_t)d_ptr) & 15) { set_error("synth_fill needs a 16-byte aligned pointer"); return FCB_ERR_INVALID_ARG; }
  synth_fill_kernel<<<148 * 8, 256, 0, st>>>((uint8_t*)d_ptr, n_bytes, seed, mask, offset);
  FCB_CUDA_OK(cudaGetLastError());
  return FCB_OK;
}

}  // namespace fcb
