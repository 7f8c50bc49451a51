// fcb_im2col.cu -- lowering of thin-input convolutions (Kx*Ky*C <= 128, e.g. the C = 3 first layer) to a 1x1 convolution
// the tensor-core engine can run: for every output pixel the (ky, kx, c)-ordered window -- exactly the word sequence
// ConvolutionInputGenerator_NonSquare emits (slidingwindow.h:1302-1313), with FMPadding_nonsquare's zeros
// (streamtools.h:361-406) and the stride decimation (conv_nonsquare_top.cpp:246-259) resolved -- is written as one
// 128-byte row (zero padded beyond K).  TMA cannot do this gather (3-byte pixels); a warp builds one row per step:
// lane l gathers bytes 4l..4l+3 through L1 and stores one coalesced 128-byte row.
#include "fcb_internal.h"

namespace fcb {

__global__ void __launch_bounds__(256) im2col_rows_kernel(const Im2colParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this lane's 4 window bytes: tap coordinates and byte offset inside the input word are fixed
  int ky[4], kx[4], co[4];
  bool kv[4];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int k = 4 * lane + b;
    kv[b] = k < p.K;
    const int tap = kv[b] ? k / p.C : 0;
    ky[b] = tap / p.KX - p.PAD; kx[b] = tap % p.KX - p.PAD; co[b] = kv[b] ? k - tap * p.C : 0;
  }
  // grid = (column blocks of 8 warps x 4 pixels, output rows, images): no divisions in the pixel loop
  const int img = blockIdx.z, oy = blockIdx.y;
  const uint8_t* in = p.in + (size_t)img * p.in_img_bytes;
  uint32_t* out = reinterpret_cast<uint32_t*>(p.out) + ((size_t)img * p.OY + oy) * p.OX * 32;
  const int ox0 = (blockIdx.x * 8 + warp) * 4;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int ox = ox0 + i;
    if (ox >= p.OX) break;
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int iy = oy * p.S + ky[b], ix = ox * p.S + kx[b];
      if (kv[b] && iy >= 0 && iy < p.IY && ix >= 0 && ix < p.IX)
        word |= (uint32_t)__ldg(in + ((size_t)iy * p.IX + ix) * p.in_word_bytes + co[b]) << (8 * b);
    }
    out[(size_t)ox * 32 + lane] = word;
  }
}

int launch_im2col(const Im2colParams& p, int n_images, cudaStream_t st) {
  for (int n0 = 0; n0 < n_images; n0 += 65535) {
    Im2colParams q = p;
    const int nb = n_images - n0 < 65535 ? n_images - n0 : 65535;
    q.in = p.in + (size_t)n0 * p.in_img_bytes;
    q.out = p.out + (size_t)n0 * p.OX * p.OY * 128;
    dim3 grid((p.OX + 31) / 32, p.OY, nb);
    im2col_rows_kernel<<<grid, 256, 0, st>>>(q);
    FCB_CUDA_OK(cudaGetLastError());
  }
  return FCB_OK;
}

// ---- 1-bit activations -> {-1,+1} bytes (tensor-core form of the xnor layer) -----------------------------------------
// A thread expands 16 channels of one pixel of the (padded) frame: two input bytes in, one 16-byte store out (K = channels rounded
// up to 16, so a pixel is K / 16 such groups).  Bit b -> byte 0x01 / 0xFF: the four bits of a nibble are spread to the low bits of
// four bytes (t), and 0xFFFFFFFF - 0xFE * t leaves 0x01 where the bit was set and 0xFF where it was not (no borrow crosses a byte).
__device__ __forceinline__ uint32_t expand_nibble(uint32_t x) {
  const uint32_t t = (x | (x << 7) | (x << 14) | (x << 21)) & 0x01010101u;
  return 0xFFFFFFFFu - 0xFEu * t;
}
// p.S = pixels per output row (1, or 2: row x = the bytes of frame pixels x and x+1 back to back; past the frame's last column
// the second half is zeros -- it only ever meets zero weights or discarded columns).
__global__ void __launch_bounds__(256) expand_bits_kernel(const Im2colParams p) {
  const int img = blockIdx.z, oy = blockIdx.y;
  const int gpp = p.K >> 4, groups = gpp * p.S;  // 16-channel groups per pixel / per output row
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.OX * groups) return;
  const int ox = idx / groups, gr = idx - ox * groups, half = gr / gpp, g = gr - half * gpp;
  const int iy = oy - p.PAD, ix = ox + half - p.PAD;
  if (ox + half >= p.OX) {  // (pair rows only) beyond the padded frame
    reinterpret_cast<uint4*>(p.out)[(((size_t)img * p.OY + oy) * p.OX + ox) * groups + gr] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  uint32_t bits = 0;  // a zero-padded border bit is an ordinary 0 activation
  if (iy >= 0 && iy < p.IY && ix >= 0 && ix < p.IX) {
    const uint8_t* word = p.in + (size_t)img * p.in_img_bytes + ((size_t)iy * p.IX + ix) * p.in_word_bytes;
    bits = word[2 * g];
    if (16 * g + 8 < p.C) bits |= (uint32_t)word[2 * g + 1] << 8;
  }
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    o[j] = expand_nibble((bits >> (4 * j)) & 0xFu);
    const int c = 16 * g + 4 * j, left = p.C - c;  // channels beyond C: 0 (their weights are 0 too)
    if (left < 4) o[j] &= left <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - left)));
  }
  reinterpret_cast<uint4*>(p.out)[(((size_t)img * p.OY + oy) * p.OX + ox) * groups + gr] = make_uint4(o[0], o[1], o[2], o[3]);
}

int launch_expand_bits(const Im2colParams& p, int n_images, cudaStream_t st) {
  const int per_row = p.OX * (p.K >> 4) * p.S;
  for (int n0 = 0; n0 < n_images; n0 += 65535) {
    Im2colParams q = p;
    const int nb = n_images - n0 < 65535 ? n_images - n0 : 65535;
    q.in = p.in + (size_t)n0 * p.in_img_bytes;
    q.out = p.out + (size_t)n0 * p.OX * p.OY * p.K * p.S;
    dim3 grid((per_row + 255) / 256, p.OY, nb);
    expand_bits_kernel<<<grid, 256, 0, st>>>(q);
    FCB_CUDA_OK(cudaGetLastError());
  }
  return FCB_OK;
}

}  // namespace fcb
