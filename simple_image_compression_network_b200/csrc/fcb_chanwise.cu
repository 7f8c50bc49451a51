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

// ------------------------------------------------------------------------------------------------------------------------
// Byte-lane form of the same units (8-bit input lanes, 8 / 16 / 32-bit output lanes, channel count a multiple of 4, words without
// container padding): a thread owns 4*V consecutive channels of one output pixel and moves them as one 4*V-byte load per tap and one
// vector store, so a warp reads 128*V consecutive bytes per tap -- these units are HBM-bound (K2 operations per byte) and what they
// need is bytes in flight, not arithmetic: 2048 threads x 16 B per SM cover the ~23 B/clk/SM that HBM delivers.  Same functions,
// same order of operations as chanwise_kernel, in 32-bit arithmetic (launch_chanwise sends here only what is exact in 32 bits:
// accumulator types of at most 31 bits for the pool functions, Kx*Ky <= 128 for the depth-wise sum of 8-bit x 16-bit products).
__device__ __forceinline__ int32_t wrap32(int32_t v, int bits, int sgn) {
  if (bits >= 32) return v;
  const uint32_t u = (uint32_t)v << (32 - bits);
  return sgn ? ((int32_t)u >> (32 - bits)) : (int32_t)(u >> (32 - bits));
}

// BM: 0 depth-wise MAC (3: the same through IDP.4A, weights of at most 8 bits), 1 max (the tap converts to the function's type without changing value: checked by the launcher), 2 the three
// sums (Avg / Acc / QuantAvg: converting every tap and wrapping every += is congruent mod 2^TA to wrapping the exact sum once).
// INS: signed input lanes.  Both are template parameters: these kernels are bound by instruction issue once the loads are wide.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {  // generic mode: selector bit 3 replicates the byte's sign
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

template <bool AS>  // d = c + sum of four (8-bit lanes of a, signed iff AS) x (signed 8-bit lanes of b)
__device__ __forceinline__ int32_t dp4a_mixed(uint32_t a, uint32_t b, int32_t c) {
  int32_t d;
  if (AS) asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Depth-wise with weights of at most 8 bits: the four taps of a chunk go through the dot-product unit.  Per word of four channels the
// 4x4 bytes (tap x channel) are transposed with 8 PRMTs into one register per channel holding its four taps, and one IDP.4A per
// channel multiplies them with that channel's four weights (wv: the [chunk][channel] words of table wt4, zero past the last tap):
// 12 instructions per 16 MACs instead of ~70.
template <int V, bool INS>
__device__ __forceinline__ void chanwise_dw_chunk(const uint32_t (&w)[4][V], const uint4 (&wv)[V], int32_t (&acc)[4 * V]) {
#pragma unroll
  for (int i = 0; i < V; i++) {
    const uint32_t lo01 = prmt(w[0][i], w[1][i], 0x5140u), hi01 = prmt(w[0][i], w[1][i], 0x7362u);
    const uint32_t lo23 = prmt(w[2][i], w[3][i], 0x5140u), hi23 = prmt(w[2][i], w[3][i], 0x7362u);
    acc[4 * i] = dp4a_mixed<INS>(prmt(lo01, lo23, 0x5410u), wv[i].x, acc[4 * i]);
    acc[4 * i + 1] = dp4a_mixed<INS>(prmt(lo01, lo23, 0x7632u), wv[i].y, acc[4 * i + 1]);
    acc[4 * i + 2] = dp4a_mixed<INS>(prmt(hi01, hi23, 0x5410u), wv[i].z, acc[4 * i + 2]);
    acc[4 * i + 3] = dp4a_mixed<INS>(prmt(hi01, hi23, 0x7632u), wv[i].w, acc[4 * i + 3]);
  }
}

// One chunk of up to four taps (cnt valid ones, t0 = index of the first) into the accumulators.  w[u][i]: word i (4 channels) of tap u,
// zeros for FMPadding and past the last tap.
template <int V, int BM, bool INS>
__device__ __forceinline__ void chanwise_consume(const ChanParams& p, const uint32_t (&w)[4][V], int t0, int cnt, int ch0, int32_t (&acc)[4 * V],
                                                 uint32_t (&mx)[2 * V]) {
  if (BM == 3) {
    uint4 wv[V];
#pragma unroll
    for (int i = 0; i < V; i++) wv[i] = __ldg(reinterpret_cast<const uint4*>(p.wt4 + (size_t)(t0 >> 2) * p.Cpad + ch0) + i);
    chanwise_dw_chunk<V, INS>(w, wv, acc);
  } else {
#pragma unroll
  for (int u = 0; u < 4; u++) {
    if (u < cnt) {
      if (BM == 2) {
        // sums: two channels per register as 16-bit halves, plain 32-bit adds (<= 128 taps x 255 < 2^16: no carry between the halves);
        // signed lanes are summed as a + 128 (padding zeros included: 0x80 after the flip) and corrected once at the end
#pragma unroll
        for (int i = 0; i < V; i++) {
          const uint32_t b = INS ? (w[u][i] ^ 0x80808080u) : w[u][i];
          mx[2 * i] += prmt(b, 0u, 0x4140u);
          mx[2 * i + 1] += prmt(b, 0u, 0x4342u);
        }
      } else if (BM == 1) {
#pragma unroll
        for (int i = 0; i < V; i++) {
          mx[2 * i] = __vmaxs2(mx[2 * i], prmt(w[u][i], 0u, INS ? 0x9180u : 0x4140u));          // bytes 0, 1 sign- / zero-extended to 16 bits
          mx[2 * i + 1] = __vmaxs2(mx[2 * i + 1], prmt(w[u][i], 0u, INS ? 0xB3A2u : 0x4342u));  // bytes 2, 3
        }
      } else {
#pragma unroll
        for (int i = 0; i < V; i++) {
          const uint2 wq = __ldg(reinterpret_cast<const uint2*>(p.wt + (size_t)(t0 + u) * p.Cpad + ch0) + i);  // 4 int16 weights
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t sh = w[u][i] >> (8 * j);
            const int32_t a = INS ? (int32_t)(int8_t)sh : (int32_t)(sh & 0xFFu);
            const uint32_t pair = (j & 2) ? wq.y : wq.x;
            acc[4 * i + j] += ((j & 1) ? ((int32_t)pair >> 16) : (int32_t)(int16_t)(pair & 0xFFFFu)) * a;
          }
        }
      }
    }
  }
  }
}

template <int V>
__device__ __forceinline__ void chanwise_load(const uint8_t* src, bool ok, uint32_t (&w)[V]) {
  if (V == 4) {
    uint4 q = make_uint4(0u, 0u, 0u, 0u);  // FMPadding zero
    if (ok) q = __ldg(reinterpret_cast<const uint4*>(src));
    w[0] = q.x; w[1 % V] = q.y; w[2 % V] = q.z; w[3 % V] = q.w;
  } else {
    w[0] = ok ? __ldg(reinterpret_cast<const uint32_t*>(src)) : 0u;
  }
}

// KS: 2 / 3 = square window of that size with unit dilation, fully unrolled (all loads of the window issued first, tap addresses from
// one row pointer per ky); 0 = any window, taps in (ky, kx) order four at a time.
// accumulators -> activation / pool function result -> packed lanes -> one vector store (pixel `pix` of image blockIdx.y)
template <int V, int BM, bool INS>
__device__ __forceinline__ void chanwise_finish(const ChanParams& p, int32_t (&acc)[4 * V], uint32_t (&mx)[2 * V], int taps, int ch0, int pix) {
  constexpr int N = 4 * V;
  if (BM == 1) {
#pragma unroll
    for (int j = 0; j < N; j++) acc[j] = (j & 1) ? ((int32_t)mx[j >> 1] >> 16) : (int32_t)(int16_t)(mx[j >> 1] & 0xFFFFu);
  } else if (BM == 2) {
    const int32_t bias = INS ? 128 * taps : 0;
#pragma unroll
    for (int j = 0; j < N; j++) acc[j] = (int32_t)((j & 1) ? (mx[j >> 1] >> 16) : (mx[j >> 1] & 0xFFFFu)) - bias;
  }
  uint32_t r[N];
  const uint32_t omask = p.out_bits >= 32 ? 0xffffffffu : ((1u << p.out_bits) - 1u);
  // accu / size (pool.hpp:151-154, C++ truncation) by multiply-high: |accu| <= 128 taps x 255 < 2^16 and size < 2^16, so
  // floor(n / d) = umulhi(n, ceil(2^32 / d)) exactly (the error term n * e / 2^32 < 2^-16 < 1 / d); one division per thread for the constant
  // the sum only needs the wrap to TA when taps x 255 can leave TA's range (a uniform branch instead of shifts on every lane)
  const bool wrap_sum = BM == 2 && ((long long)taps * 255 >= (1ll << (p.acc_bits - (p.acc_signed ? 1 : 0))) || (INS && !p.acc_signed));  // (negative sums in an unsigned TA wrap too)
  const uint32_t div_m = (BM == 2 && p.mode == CW_POOL_AVG && p.size > 1) ? 0xFFFFFFFFu / (uint32_t)p.size + 1u : 0u;
  if ((BM == 0 || BM == 3) && p.epi.act_kind == FCB_ACT_PASSTHROUGH) {
    // PassThroughActivation (activations.hpp:127-134): the TA-wrapped sum, truncated to the output lane -- uniform branches instead of
    // the generic activate() per channel (the finish is a third of this kernel's instructions otherwise)
    if (p.epi.acc_bits < 32) {
#pragma unroll
      for (int j = 0; j < N; j++) r[j] = (uint32_t)wrap_ta(acc[j], p.epi.acc_bits, p.epi.acc_signed) & omask;
    } else {
#pragma unroll
      for (int j = 0; j < N; j++) r[j] = (uint32_t)acc[j] & omask;
    }
  } else if ((BM == 0 || BM == 3) && p.epi.act_kind == FCB_ACT_THRESHOLDS) {
    // ThresholdsActivation: the N channels' binary searches advance in lock step, so the N table loads of a level are in flight
    // together (activate() per channel would chain N x log2(thr_n + 1) dependent loads)
    const EpiParams& e = p.epi;
    int32_t a[N];
    int pos[N];
#pragma unroll
    for (int j = 0; j < N; j++) { a[j] = wrap_ta(acc[j], e.acc_bits, e.acc_signed); pos[j] = 0; }
    const bool strict = (e.cmp == FCB_CMP_LESS) || (e.cmp == FCB_CMP_GREATER_EQUAL);
    const int32_t* __restrict__ t = e.thr + ch0;
#pragma unroll 1
    for (int step = (e.thr_n + 1) >> 1; step; step >>= 1) {
      int32_t tv[N];
#pragma unroll
      for (int j = 0; j < N; j++) tv[j] = __ldg(t + (size_t)(pos[j] + step - 1) * e.thr_stride + j);
#pragma unroll
      for (int j = 0; j < N; j++) pos[j] += (strict ? (tv[j] < a[j]) : (tv[j] <= a[j])) ? step : 0;
    }
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = thr_finish(e, pos[j]);
  } else
#pragma unroll
  for (int j = 0; j < N; j++) {
    if (BM == 0 || BM == 3) {
      r[j] = activate(p.epi, ch0 + j, acc[j]);
    } else {
      int32_t o = acc[j];
      if (BM == 2) {
        if (wrap_sum) o = wrap32(o, p.acc_bits, p.acc_signed);
        if (p.mode == CW_POOL_AVG) {
          if (p.size > 1) {
            const int32_t q = (int32_t)__umulhi((uint32_t)abs(o), div_m);
            o = o < 0 ? -q : q;
          } else if (p.size == 0) {
            o = 0;
          }
        }
        else if (p.mode == CW_POOL_QUANTAVG) o = o >> p.size;
      }
      r[j] = (uint32_t)o & omask;
    }
  }
  uint8_t* dst = p.out + (size_t)blockIdx.y * p.out_img_bytes + (size_t)pix * p.out_word_bytes + (size_t)ch0 * (p.out_bits >> 3);
  if (p.out_bits == 8) {
    uint32_t o[V];
#pragma unroll
    for (int i = 0; i < V; i++) o[i] = (r[4 * i] & 0xFFu) | ((r[4 * i + 1] & 0xFFu) << 8) | ((r[4 * i + 2] & 0xFFu) << 16) | (r[4 * i + 3] << 24);
    if (V == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1 % V], o[2 % V], o[3 % V]);
    else *reinterpret_cast<uint32_t*>(dst) = o[0];
  } else if (p.out_bits == 16) {
    uint32_t o[2 * V];
#pragma unroll
    for (int i = 0; i < 2 * V; i++) o[i] = (r[2 * i] & 0xFFFFu) | (r[2 * i + 1] << 16);
    if (V == 4) {
#pragma unroll
      for (int i = 0; i < V / 2; i++) reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[4 * i], o[(4 * i + 1) % (2 * V)], o[(4 * i + 2) % (2 * V)], o[(4 * i + 3) % (2 * V)]);
    } else {
      *reinterpret_cast<uint2*>(dst) = make_uint2(o[0], o[1]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; i++) reinterpret_cast<uint4*>(dst)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
  }
}

// KS: 2 / 3 = square window of that size with unit dilation, fully unrolled (all loads of the window issued first, tap addresses from
// one row pointer per ky); 0 = any window, taps in (ky, kx) order four at a time.  VO = 2 (3x3 windows at vertical stride 1): a thread
// owns two vertically adjacent outputs and loads the four input rows they share once (12 loads for two outputs instead of 18).
template <int V, int BM, bool INS, int KS, int VO>
__global__ void __launch_bounds__(256) chanwise_bytes_kernel(const ChanParams p) {
  constexpr int N = 4 * V;  // channels per thread
  const int groups = p.C / N, oyb = (p.OY + VO - 1) / VO;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)p.OX * oyb * groups) return;
  const int g = (int)(t % groups), pix = (int)(t / groups);
  const int oy = (pix / p.OX) * VO, ox = pix - (pix / p.OX) * p.OX, ch0 = g * N;
  const uint8_t* in = p.in + (size_t)blockIdx.y * p.in_img_bytes + (size_t)ch0;
  const int y0 = oy * p.SY - p.pad_u, x0 = ox * p.SX - p.pad_l;
  int32_t acc[VO][N];
  uint32_t mx[VO][N / 2];  // max and sums: two channels per register as 16-bit halves (VIMNMX.S16x2 / one IADD; byte lanes widen with one PRMT per pair)
  {
    // pool.hpp:98-102: the type's minimum, or min_value (maxpool.h:144-150); anything below -32768 is below every 8-bit tap
    int32_t first = BM == 1 ? (p.has_init ? p.init : (p.acc_signed ? -(1 << (p.acc_bits - 1)) : 0)) : 0;
    first = max(first, -32768);
#pragma unroll
    for (int vo = 0; vo < VO; vo++) {
#pragma unroll
      for (int j = 0; j < N; j++) acc[vo][j] = 0;
#pragma unroll
      for (int j = 0; j < N / 2; j++) mx[vo][j] = BM == 1 ? ((uint32_t)first & 0xFFFFu) * 0x10001u : 0u;
    }
  }
  const int taps = p.KX * p.KY;
  if constexpr (KS > 0) {
    constexpr int T = KS * KS, ROWS = KS + VO - 1, TL = ROWS * KS, TC = KS ? (TL + 3) / 4 * 4 + 4 : 4;  // (KS = 0 never runs this branch)
    uint32_t w[TC][V];
#pragma unroll
    for (int ky = 0; ky < ROWS; ky++) {
      const int y = y0 + ky;
      const bool yok = y >= 0 && y < p.IY;
      const uint8_t* row = in + ((size_t)(yok ? y : 0) * p.IX) * p.in_word_bytes;
#pragma unroll
      for (int kx = 0; kx < KS; kx++) {
        const int x = x0 + kx;
        const bool ok = yok && x >= 0 && x < p.IX;
        chanwise_load<V>(row + (size_t)(ok ? x : 0) * p.in_word_bytes, ok, w[ky * KS + kx]);
      }
    }
#pragma unroll
    for (int t0 = 0; t0 < T; t0 += 4) {
      uint4 wv[V];  // (depth-wise: the chunk's weights are loaded once for all the thread's outputs)
      if (BM == 3) {
#pragma unroll
        for (int i = 0; i < V; i++) wv[i] = __ldg(reinterpret_cast<const uint4*>(p.wt4 + (size_t)(t0 >> 2) * p.Cpad + ch0) + i);
      }
#pragma unroll
      for (int vo = 0; vo < VO; vo++) {
        uint32_t wc[4][V];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int i = 0; i < V; i++) wc[u][i] = t0 + u < T ? w[vo * KS + t0 + u][i] : 0u;  // output vo starts KS taps (one row) further
        if (BM == 3) chanwise_dw_chunk<V, INS>(wc, wv, acc[vo]);
        else chanwise_consume<V, BM, INS>(p, wc, t0, T - t0 < 4 ? T - t0 : 4, ch0, acc[vo], mx[vo]);
      }
    }
  } else {
    // taps in (ky, kx) order, four at a time: the loads of a chunk are issued together (what these units need is bytes in flight)
    int ky = 0, kx = 0;
    for (int t0 = 0; t0 < taps; t0 += 4) {
      uint32_t w[4][V];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int y = y0 + ky * p.DY, x = x0 + kx * p.DX;
        const bool ok = t0 + u < taps && y >= 0 && y < p.IY && x >= 0 && x < p.IX;
        chanwise_load<V>(in + ((size_t)(ok ? y : 0) * p.IX + (ok ? x : 0)) * p.in_word_bytes, ok, w[u]);
        if (++kx == p.KX) { kx = 0; ++ky; }
      }
      chanwise_consume<V, BM, INS>(p, w, t0, taps - t0 < 4 ? taps - t0 : 4, ch0, acc[0], mx[0]);
    }
  }
#pragma unroll
  for (int vo = 0; vo < VO; vo++)
    if (oy + vo < p.OY) chanwise_finish<V, BM, INS>(p, acc[vo], mx[vo], taps, ch0, (oy + vo) * p.OX + ox);
}

template <int V, int KS, int VO>
static void launch_bytes_k(const ChanParams& q, int nb, cudaStream_t st) {
  const long long threads = (long long)q.OX * ((q.OY + VO - 1) / VO) * (q.C / (4 * V));
  const dim3 grid((unsigned)((threads + 255) / 256), nb, 1);
  const int bm = q.mode == CW_DWCONV ? (q.wt4 ? 3 : 0) : q.mode == CW_POOL_MAX ? 1 : 2;
  if (q.in_signed) {
    if (bm == 1) chanwise_bytes_kernel<V, 1, true, KS, VO><<<grid, 256, 0, st>>>(q);
    else if (bm == 2) chanwise_bytes_kernel<V, 2, true, KS, VO><<<grid, 256, 0, st>>>(q);
    else chanwise_bytes_kernel<V, 3, true, KS, VO><<<grid, 256, 0, st>>>(q);
  } else {
    if (bm == 1) chanwise_bytes_kernel<V, 1, false, KS, VO><<<grid, 256, 0, st>>>(q);
    else if (bm == 2) chanwise_bytes_kernel<V, 2, false, KS, VO><<<grid, 256, 0, st>>>(q);
    else chanwise_bytes_kernel<V, 3, false, KS, VO><<<grid, 256, 0, st>>>(q);
  }
}
template <int V>
static void launch_bytes(const ChanParams& q, int nb, cudaStream_t st) {
  const int ks = (q.KX == q.KY && q.DX == 1 && q.DY == 1 && (q.KX == 2 || q.KX == 3)) ? q.KX : 0;
  if (q.mode == CW_DWCONV && !q.wt4) {  // 16-bit weights: the rolled form keeps the registers down
    const long long threads = (long long)q.OX * q.OY * (q.C / (4 * V));
    const dim3 grid((unsigned)((threads + 255) / 256), nb, 1);
    if (q.in_signed) chanwise_bytes_kernel<V, 0, true, 0, 1><<<grid, 256, 0, st>>>(q);
    else chanwise_bytes_kernel<V, 0, false, 0, 1><<<grid, 256, 0, st>>>(q);
  } else if (ks == 3 && q.SY == 1 && q.OY > 1) launch_bytes_k<V, 3, 2>(q, nb, st);
  else if (ks == 3) launch_bytes_k<V, 3, 1>(q, nb, st);
  else if (ks == 2) launch_bytes_k<V, 2, 1>(q, nb, st);
  else launch_bytes_k<V, 0, 1>(q, nb, st);
}

// AddStreams_Batch on byte lanes (8-bit operands, 8- or 16-bit sums): 16 lanes per thread
__global__ void __launch_bounds__(256) add_streams_bytes_kernel(const uint4* __restrict__ in1, const uint4* __restrict__ in2, uint8_t* __restrict__ out,
                                                                unsigned long long n16, int s1, int s2, int ob, int offset) {
  const unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= n16) return;
  const uint4 a = __ldg(in1 + t), b = __ldg(in2 + t);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  uint32_t r[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const uint32_t x = (aw[j >> 2] >> (8 * (j & 3))) & 0xFFu, y = (bw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
    r[j] = (uint32_t)((s1 ? (int32_t)(int8_t)x : (int32_t)x) + (s2 ? (int32_t)(int8_t)y : (int32_t)y) + offset);
  }
  if (ob == 8) {
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = (r[4 * i] & 0xFFu) | ((r[4 * i + 1] & 0xFFu) << 8) | ((r[4 * i + 2] & 0xFFu) << 16) | (r[4 * i + 3] << 24);
    reinterpret_cast<uint4*>(out)[t] = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 2; i++)
      reinterpret_cast<uint4*>(out)[2 * t + i] = make_uint4((r[8 * i] & 0xFFFFu) | (r[8 * i + 1] << 16), (r[8 * i + 2] & 0xFFFFu) | (r[8 * i + 3] << 16),
                                                           (r[8 * i + 4] & 0xFFFFu) | (r[8 * i + 5] << 16), (r[8 * i + 6] & 0xFFFFu) | (r[8 * i + 7] << 16));
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
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (b1 == 8 && b2 == 8 && (ob == 8 || ob == 16) && wb1 == ch && wb2 == ch && wbo == ch * ob / 8 && ((unsigned long long)ch * n_words) % 16 == 0 &&
      al16(d_in1) && al16(d_in2) && al16(d_out)) {
    // byte lanes without container padding: the word structure does not matter, the stream is a flat array of lanes
    const unsigned long long n16 = (unsigned long long)ch * n_words / 16, blocks = (n16 + 255) / 256;
    for (unsigned long long b0 = 0; b0 < blocks; b0 += 0x40000000ull) {
      const unsigned nb = (unsigned)std::min<unsigned long long>(0x40000000ull, blocks - b0);
      add_streams_bytes_kernel<<<nb, 256, 0, st>>>((const uint4*)d_in1 + b0 * 256, (const uint4*)d_in2 + b0 * 256, (uint8_t*)d_out + b0 * 256 * 2 * ob,
                                                   n16 - b0 * 256, s1, s2, ob, offset);
      FCB_CUDA_OK(cudaGetLastError());
    }
    return FCB_OK;
  }
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

// 32-bit words of channels a thread of chanwise_bytes_kernel would own for this unit (4 or 1), 0 = the general kernel: what the
// byte-lane form computes exactly in 32 bits, on words it can move as vectors (buffer alignment is checked again per launch)
int chanwise_vector_words(const ChanParams& p) {
  // depth-wise / sums: exact in 32 bits (8-bit lanes x 16-bit weights x <= 128 taps); max: the tap must convert to the function's
  // type without changing value (same signedness and >= 8 bits, or unsigned lanes in a signed type of >= 9 bits)
  const bool exact32 = p.KX * p.KY <= 128 &&
                       (p.mode == CW_DWCONV ? true
                        : p.mode == CW_POOL_MAX
                            ? (p.acc_bits <= 31 && (!p.has_init || p.init <= 32767) &&
                               ((p.acc_signed == p.in_signed && p.acc_bits >= 8) || (p.acc_signed && !p.in_signed && p.acc_bits >= 9)))
                            : (p.acc_bits <= 31 && p.size >= 0 && p.size < 32));
  if (!(exact32 && p.in_bits == 8 && (p.out_bits == 8 || p.out_bits == 16 || p.out_bits == 32) && p.C % 4 == 0 && p.in_word_bytes == p.C &&
        p.out_word_bytes == p.C * (p.out_bits / 8) && (p.mode != CW_DWCONV || p.Cpad % 4 == 0) && p.in_img_bytes % 16 == 0 &&
        p.out_img_bytes % 16 == 0))
    return 0;
  return p.C % 16 == 0 ? 4 : 1;
}

int launch_chanwise(const ChanParams& p, int n_images, cudaStream_t st) {
  const int blocks = (p.OX * p.OY + 7) / 8;
  int V = chanwise_vector_words(p);
  if (((reinterpret_cast<uintptr_t>(p.in) | reinterpret_cast<uintptr_t>(p.out)) & 15) ||
      (p.mode == CW_DWCONV && ((reinterpret_cast<uintptr_t>(p.wt) & 7) || (reinterpret_cast<uintptr_t>(p.wt4) & 15))))
    V = 0;
  for (int n0 = 0; n0 < n_images; n0 += 65535) {
    ChanParams q = p;
    const int nb = n_images - n0 < 65535 ? n_images - n0 : 65535;
    q.in = p.in + (size_t)n0 * p.in_img_bytes;
    q.out = p.out + (size_t)n0 * p.out_img_bytes;
    if (V) {
      if (V == 4) launch_bytes<4>(q, nb, st);
      else launch_bytes<1>(q, nb, st);
    } else
    chanwise_kernel<<<dim3(blocks, nb, 1), 256, 0, st>>>(q);
    FCB_CUDA_OK(cudaGetLastError());
  }
  return FCB_OK;
}

}  // namespace fcb
