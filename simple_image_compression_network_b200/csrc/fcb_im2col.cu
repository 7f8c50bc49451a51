// fcb_im2col.cu -- lowering of thin-input convolutions (Kx*Ky*C <= 128, e.g. the C = 3 first layer) to a 1x1 convolution
// the tensor-core engine can run: for every output pixel the (ky, kx, c)-ordered window -- exactly the word sequence
// ConvolutionInputGenerator_NonSquare emits (slidingwindow.h:1302-1313), with FMPadding_nonsquare's zeros
// (streamtools.h:361-406) and the stride decimation (conv_nonsquare_top.cpp:246-259) resolved -- is written as one
// 128-byte row (zero padded beyond K).  TMA cannot do this gather (3-byte pixels); a warp builds one row per step:
// lane l gathers bytes 4l..4l+3 through L1 and stores one coalesced 128-byte row.
#include "fcb_internal.h"

namespace fcb {

__global__ void __launch_bounds__(256) im2col_rows_kernel(const Im2colParams p, int n_images) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long pix_per_img = (long long)p.OX * p.OY, total = pix_per_img * n_images;
  // this lane's 4 window bytes: tap coordinates and byte offset inside the input word are fixed
  int ky[4], kx[4], co[4];
  bool kv[4];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int k = 4 * lane + b;
    kv[b] = k < p.K;
    const int tap = kv[b] ? k / p.C : 0;
    ky[b] = tap / p.KX; kx[b] = tap % p.KX; co[b] = kv[b] ? k - tap * p.C : 0;
  }
  for (long long w = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < total; w += warps) {
    const int img = (int)(w / pix_per_img);
    const int pix = (int)(w - (long long)img * pix_per_img);
    const int oy = pix / p.OX, ox = pix - oy * p.OX;
    const uint8_t* in = p.in + (size_t)img * p.in_img_bytes;
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int iy = oy * p.S + ky[b] - p.PAD, ix = ox * p.S + kx[b] - p.PAD;
      if (kv[b] && iy >= 0 && iy < p.IY && ix >= 0 && ix < p.IX)
        word |= (uint32_t)__ldg(in + ((size_t)iy * p.IX + ix) * p.in_word_bytes + co[b]) << (8 * b);
    }
    reinterpret_cast<uint32_t*>(p.out)[w * 32 + lane] = word;
  }
}

int launch_im2col(const Im2colParams& p, int n_images, cudaStream_t st) {
  im2col_rows_kernel<<<148 * 16, 256, 0, st>>>(p, n_images);
  FCB_CUDA_OK(cudaGetLastError());
  return FCB_OK;
}

}  // namespace fcb
