// fcb_chanwise.cu -- the channel-wise units of the library: one output channel per input channel.
//
//   FCB_KIND_DWCONV : depth-wise convolution = ConvolutionInputGenerator[_NonSquare]_dws (slidingwindow.h:761-868, 1377-1488)
//                     + Vector_Vector_Activate_Batch (vvau.hpp:80-154): acc[ch] = sum_k W[ch][k] * a_k[ch], k = ky*Kx + kx (the
//                     generator emits, per channel chunk, the taps in (ky, kx) order), then the activation stage (fcb_epilogue.cuh).
//   FCB_KIND_POOL   : the same sliding window + Pool_batch (maxpool.h:525-577) with MaxPoolFunction / AvgPoolFunction /
//                     AccPoolFunction / QuantAvgPoolFunction (pool.hpp:94-226): init(), pool() per tap in the function's type,
//                     activate() once.
//   FMPadding_nonsquare (streamtools.h:361-406) in front is resolved while reading: out-of-frame taps are zeros.
//
// These are streaming units (K2 MACs or compares per output byte): a warp owns one output pixel, its lanes walk the channels,
// so loads and stores of byte lanes are 32 consecutive bytes of one stream word; taps of neighbouring pixels hit L1/L2.
#include "fcb_epilogue.cuh"

namespace fcb {

__device__ __forceinline__ int64_t wrap64(int64_t v, int bits, int sgn) {
  if (bits >= 64) return v;
  const uint64_t u = (uint64_t)v << (64 - bits);
  return sgn ? ((int64_t)u >> (64 - bits)) : (int64_t)(u >> (64 - bits));
}

__global__ void __launch_bounds__(256) chanwise_kernel(const ChanParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix = blockIdx.x * 8 + warp;
  if (pix >= p.OX * p.OY) return;  // (warp-uniform)
  const int oy = pix / p.OX, ox = pix - oy * p.OX;
  const uint8_t* in = p.in + (size_t)blockIdx.y * p.in_img_bytes;
  uint8_t* oword = p.out + (size_t)blockIdx.y * p.out_img_bytes + (size_t)pix * p.out_word_bytes;
  const int y0 = oy * p.SY - p.pad_u, x0 = ox * p.SX - p.pad_l;
  for (int c0 = 0; c0 < p.C; c0 += 32) {
    const int ch = c0 + lane;
    const bool chv = ch < p.C;
    int64_t acc = 0;
    if (p.mode == CW_POOL_MAX)  // pool.hpp:98-102: the type's minimum; StreamingMaxPool_Precision: min_value (maxpool.h:144-150)
      acc = p.has_init ? (int64_t)p.init : (p.acc_signed ? -((int64_t)1 << (p.acc_bits - 1)) : 0);
    for (int ky = 0; ky < p.KY; ky++) {
      const int y = y0 + ky;
      for (int kx = 0; kx < p.KX; kx++) {
        const int x = x0 + kx;
        int32_t a = 0;  // FMPadding zero
        if (chv && y >= 0 && y < p.IY && x >= 0 && x < p.IX)
          a = load_lane_any(in + ((size_t)y * p.IX + x) * p.in_word_bytes, ch, p.in_bits, p.in_signed);
        if (p.mode == CW_DWCONV) {
          const int32_t w = chv ? (int32_t)__ldg(p.wt + (size_t)(ky * p.KX + kx) * p.Cpad + ch) : 0;
          acc += (int64_t)w * a;  // exact in 64 bits; wrapping to TA once at the end is identical to wrapping at every += (mod 2^TA)
        } else {
          const int64_t v = wrap64(a, p.acc_bits, p.acc_signed);  // the slice converts to the function's type
          if (p.mode == CW_POOL_MAX) acc = v > acc ? v : acc;
          else acc = wrap64(acc + v, p.acc_bits, p.acc_signed);
        }
      }
    }
    uint32_t r;
    if (p.mode == CW_DWCONV) {
      r = chv ? activate(p.epi, ch, (int32_t)wrap64(acc, 32, 1)) : 0u;  // TA <= 32 bits: the low 32 bits determine the wrapped value
    } else {
      int64_t o = acc;
      if (p.mode == CW_POOL_AVG) o = p.size ? acc / (int64_t)p.size : 0;  // accu / size: C++ truncation (pool.hpp:151-154)
      else if (p.mode == CW_POOL_QUANTAVG) o = acc >> p.size;             // TO(accu >> size) (pool.hpp:221-224)
      r = (uint32_t)((uint64_t)o & (p.out_bits >= 32 ? 0xffffffffull : ((1ull << p.out_bits) - 1ull)));
    }
    store_lane(oword, ch, chv, r, p.out_bits);
  }
}

int launch_chanwise(const ChanParams& p, int n_images, cudaStream_t st) {
  const int blocks = (p.OX * p.OY + 7) / 8;
  for (int n0 = 0; n0 < n_images; n0 += 65535) {
    ChanParams q = p;
    const int nb = n_images - n0 < 65535 ? n_images - n0 : 65535;
    q.in = p.in + (size_t)n0 * p.in_img_bytes;
    q.out = p.out + (size_t)n0 * p.out_img_bytes;
    chanwise_kernel<<<dim3(blocks, nb, 1), 256, 0, st>>>(q);
    FCB_CUDA_OK(cudaGetLastError());
  }
  return FCB_OK;
}

}  // namespace fcb
