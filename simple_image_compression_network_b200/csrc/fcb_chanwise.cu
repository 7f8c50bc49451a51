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
#include <algorithm>

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
      const int y = y0 + ky * p.DY;
      for (int kx = 0; kx < p.KX; kx++) {
        const int x = x0 + kx * p.DX;
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

// AddStreams_Batch (streamtools.h:669-720): Out_t sum = op1 + op2 + offset per lane; a warp owns one stream word
__device__ __forceinline__ int32_t load_lane32(const uint8_t* word, int c, int bits, int sgn) {  // lanes up to 32 bits
  const size_t bit = (size_t)c * bits;
  const uint8_t* b = word + (bit >> 3);
  const int sh = (int)(bit & 7), need = (sh + bits + 7) >> 3;
  uint64_t v = 0;
  for (int i = 0; i < need; i++) v |= (uint64_t)b[i] << (8 * i);
  v >>= sh;
  if (bits < 32) v &= (1ull << bits) - 1ull;
  if (sgn && bits < 32) {
    const uint64_t m = 1ull << (bits - 1);
    v = (v ^ m) - m;
  }
  return (int32_t)(uint32_t)v;
}
__global__ void __launch_bounds__(256) add_streams_kernel(const uint8_t* __restrict__ in1, const uint8_t* __restrict__ in2, uint8_t* __restrict__ out,
                                                          unsigned long long n_words, int ch, int b1, int s1, int b2, int s2, int ob, int offset,
                                                          int wb1, int wb2, int wbo) {
  const int lane = threadIdx.x & 31;
  const unsigned long long w = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_words) return;
  for (int c0 = 0; c0 < ch; c0 += 32) {
    const int c = c0 + lane;
    const bool v = c < ch;
    uint32_t r = 0;
    if (v) r = (uint32_t)(load_lane32(in1 + w * wb1, c, b1, s1) + load_lane32(in2 + w * wb2, c, b2, s2) + offset);  // mod 2^32, then Out_t's width
    store_lane(out + w * wbo, c, v, ob >= 32 ? r : (r & ((1u << ob) - 1u)), ob);
  }
}
int launch_add_streams(const void* d_in1, const void* d_in2, void* d_out, unsigned long long n_words, int ch, int b1, int s1, int b2, int s2, int ob,
                       int offset, int wb1, int wb2, int wbo, cudaStream_t st) {
  if (!n_words) return FCB_OK;
  if ((size_t)wbo * 8 != (size_t)ch * ob) FCB_CUDA_OK(cudaMemsetAsync(d_out, 0, (size_t)wbo * n_words, st));  // container padding bits stay zero
  const unsigned long long blocks = (n_words + 7) / 8;
  for (unsigned long long b0 = 0; b0 < blocks; b0 += 0x40000000ull) {
    const unsigned nb = (unsigned)std::min<unsigned long long>(0x40000000ull, blocks - b0);
    add_streams_kernel<<<nb, 256, 0, st>>>((const uint8_t*)d_in1 + b0 * 8 * wb1, (const uint8_t*)d_in2 + b0 * 8 * wb2, (uint8_t*)d_out + b0 * 8 * wbo,
                                           n_words - b0 * 8, ch, b1, s1, b2, s2, ob, offset, wb1, wb2, wbo);
    FCB_CUDA_OK(cudaGetLastError());
  }
  return FCB_OK;
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
