// fcb_epilogue.cuh -- the activation stage shared by every engine.
//
// Restates, on an int32 accumulator:
//   TA wrap                  mvau.hpp:112, mac.hpp:166-169 (every += wraps to TA; reducing once is identical)
//   PassThroughActivation    activations.hpp:127-134 + lane truncation at mvau.hpp:167
//   bias + ReLU(wrap)        conv_nonsquare_top.cpp:267-278  ((lane + bias) mod 2^B, MSB set -> 0)
//   ThresholdsActivation     activations.hpp:168-190 with comp::{less,greater,less_equal,greater_equal} (:57-99)
// Thresholds are pre-sorted per channel at create time (the reference result is a count, so the
// order is irrelevant) which turns the NumTH compares into one binary search.
#pragma once
#include "fcb_internal.h"

namespace fcb {

// Lane `c` of a packed stream word: bits [c*bits, (c+1)*bits), lane 0 at the LSB (interpret.hpp:191-217), any width 1..16,
// reinterpreted as signed through the ap_int<bits> cast when `sgn` (interpret.hpp:213-217).
__device__ __forceinline__ int32_t load_lane_any(const uint8_t* word, int c, int bits, int sgn) {
  const size_t bit = (size_t)c * bits;
  const uint8_t* b = word + (bit >> 3);
  const int sh = (int)(bit & 7);
  uint32_t v = b[0];
  if (sh + bits > 8) v |= (uint32_t)b[1] << 8;
  if (sh + bits > 16) v |= (uint32_t)b[2] << 16;
  v = (v >> sh) & ((1u << bits) - 1u);
  if (sgn) {
    const uint32_t m = 1u << (bits - 1);
    return (int32_t)((v ^ m) - m);
  }
  return (int32_t)v;
}

__device__ __forceinline__ int32_t wrap_ta(int32_t acc, int bits, int sgn) {
  if (bits >= 32) return acc;
  const uint32_t u = (uint32_t)acc << (32 - bits);
  return sgn ? ((int32_t)u >> (32 - bits)) : (int32_t)(u >> (32 - bits));
}

// Branch-free lower/upper bound over the per-channel sorted, INT32_MAX-padded table of thr_n = 2^k - 1 entries laid out
// threshold-major ([i][channel]): consecutive lanes = consecutive channels read one 128-byte line at the top levels.
// Returns the number of table entries that compare "below" the accumulator.
__device__ __forceinline__ int thr_search(const EpiParams& e, int ch, int32_t a) {
  const int32_t* __restrict__ t = e.thr + ch;
  const bool strict = (e.cmp == FCB_CMP_LESS) || (e.cmp == FCB_CMP_GREATER_EQUAL);
  int pos = 0;
  for (int step = (e.thr_n + 1) >> 1; step; step >>= 1) {
    const int32_t tv = __ldg(t + (size_t)(pos + step - 1) * e.thr_stride);
    pos += (strict ? (tv < a) : (tv <= a)) ? step : 0;
  }
  return pos;
}
__device__ __forceinline__ uint32_t thr_finish(const EpiParams& e, int pos) {
  const uint32_t omask = e.out_bits >= 32 ? 0xffffffffu : ((1u << e.out_bits) - 1u);
  pos = min(pos, e.num_th);  // a == INT32_MAX can "pass" the padding with the non-strict compares
  const int cnt = (e.cmp == FCB_CMP_LESS || e.cmp == FCB_CMP_LESS_EQUAL) ? pos : (e.num_th - pos);
  return (uint32_t)(e.act_val + cnt) & omask;
}

__device__ __forceinline__ uint32_t activate(const EpiParams& e, int ch, int32_t acc) {
  const int32_t a = wrap_ta(acc, e.acc_bits, e.acc_signed);
  const uint32_t omask = e.out_bits >= 32 ? 0xffffffffu : ((1u << e.out_bits) - 1u);
  if (e.act_kind == FCB_ACT_PASSTHROUGH) return (uint32_t)a & omask;
  if (e.act_kind == FCB_ACT_BIAS_RELU) {
    const uint32_t r = ((uint32_t)a + (uint32_t)(int32_t)e.bias[ch]) & omask;
    return ((r >> (e.out_bits - 1)) & 1u) ? 0u : r;
  }
  return thr_finish(e, thr_search(e, ch, a));
}

// N independent searches advanced in lock step: the N loads of one level are issued back to back, so the latency of a
// level (the table does not fit the little L1 left beside the operand planes) is paid once per N outputs.
// `tbl` / `stride` select the table: the global one (e.thr, e.thr_stride) or a shared-memory copy of a channel block.
template <int N>
__device__ __forceinline__ void activate_thrN(const EpiParams& e, const int32_t* __restrict__ tbl, int stride, const int32_t (&acc)[N],
                                              uint32_t (&out)[N]) {
  int32_t a[N];
  int pos[N];
#pragma unroll
  for (int i = 0; i < N; i++) { a[i] = wrap_ta(acc[i], e.acc_bits, e.acc_signed); pos[i] = 0; }
  const bool strict = (e.cmp == FCB_CMP_LESS) || (e.cmp == FCB_CMP_GREATER_EQUAL);
#pragma unroll 1
  for (int step = (e.thr_n + 1) >> 1; step; step >>= 1) {
    int32_t tv[N];
#pragma unroll
    for (int i = 0; i < N; i++) tv[i] = tbl[(size_t)(pos[i] + step - 1) * stride];
#pragma unroll
    for (int i = 0; i < N; i++) pos[i] += (strict ? (tv[i] < a[i]) : (tv[i] <= a[i])) ? step : 0;
  }
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = thr_finish(e, pos[i]);
}

// Hybrid search of the tensor-core epilogue (thread = one fixed channel).  The top `L` levels of the binary search only
// ever probe sorted indices j * 2^(D-L) - 1 (j = 1 .. 2^L - 1; D = log2(thr_n + 1)): those 2^L - 1 entries per channel sit
// in shared memory, threshold-major, so lane <-> bank and a probe is one conflict-free wavefront whatever index each lane
// is at.  What remains is an aligned group of G = 2^(D-L) consecutive sorted thresholds: ONE 16-byte load per lane from the
// channel-major global copy (G = 4), counted in registers.  A plain search in global memory costs one L1 tag-stage
// wavefront per lane and probe (~127 per warp and output) and was the limiter at 1 wavefront/clk/SM.
__device__ __forceinline__ int32_t lds_s32(uint32_t saddr) {
  int32_t v;
  asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
// top_saddr: shared-space byte address of this channel's column of the top table; row_shift: log2(bytes per table row);
// gshift: log2(G).  32-bit shared addressing and shifts keep a level at ~5 instructions per output.
template <int N, bool WRAP = true>  // WRAP = false: the caller already holds TA-wrapped values (pooled epilogue)
__device__ __forceinline__ void activate_thr_hybrid(const EpiParams& e, uint32_t top_saddr, int row_shift, int top_levels, int gshift,
                                                    const int32_t* __restrict__ row_cm /*global row of the channel*/,
                                                    const int32_t (&acc)[N], uint32_t (&out)[N]) {
  // The running position is kept as a BYTE offset into the thread's column of the top table (posb = (pos >> gshift) << row_shift,
  // plus the column's address), so a level is: one LDS at posb + (uniform step offset), one compare, one predicated add.
  int32_t a[N];
  uint32_t posb[N];
  const uint32_t base = top_saddr - (1u << row_shift);  // row (j - 1) holds sorted index j * G - 1
#pragma unroll
  for (int i = 0; i < N; i++) { a[i] = WRAP ? wrap_ta(acc[i], e.acc_bits, e.acc_signed) : acc[i]; posb[i] = base; }
  const bool strict = (e.cmp == FCB_CMP_LESS) || (e.cmp == FCB_CMP_GREATER_EQUAL);
  uint32_t stepb = (uint32_t)(((e.thr_n + 1) >> 1) >> gshift) << row_shift;  // warp-uniform
#pragma unroll 1
  for (int l = 0; l < top_levels; l++, stepb >>= 1) {
    int32_t tv[N];
#pragma unroll
    for (int i = 0; i < N; i++) tv[i] = lds_s32(posb[i] + stepb);
    if (strict) {
#pragma unroll
      for (int i = 0; i < N; i++)
        if (tv[i] < a[i]) posb[i] += stepb;
    } else {
#pragma unroll
      for (int i = 0; i < N; i++)
        if (tv[i] <= a[i]) posb[i] += stepb;
    }
  }
  int pos[N];
#pragma unroll
  for (int i = 0; i < N; i++) pos[i] = (int)(((posb[i] - base) >> row_shift) << gshift);
  if (gshift == 2) {  // G = 4: entries pos, pos+1, pos+2 of the sorted row decide the last two levels
    int4 qv[N];
    const char* rowb = reinterpret_cast<const char*>(row_cm);
#pragma unroll
    for (int i = 0; i < N; i++) qv[i] = __ldg(reinterpret_cast<const int4*>(rowb + (uint32_t)(((posb[i] - base) >> row_shift) << 4)));
#pragma unroll
    for (int i = 0; i < N; i++) {
      pos[i] += strict ? ((qv[i].x < a[i]) + (qv[i].y < a[i]) + (qv[i].z < a[i])) : ((qv[i].x <= a[i]) + (qv[i].y <= a[i]) + (qv[i].z <= a[i]));
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = thr_finish(e, pos[i]);
}

// Bucket-LUT search over a FULL-depth shared-memory table (threshold-major, row i = sorted index i): the LUT byte gives the sorted
// index the accumulator's bucket starts at (clamped so that the 2^levels-wide window stays inside the 2^D - 1 rows); `levels`
// binary levels resolve the <= 2^levels - 1 thresholds inside the bucket.  Thresholds of later buckets are larger than every value of this bucket, so
// probing past the bucket's end can only compare false: exact for all four comp:: functors.
template <int N>
__device__ __forceinline__ void activate_thr_lut(const EpiParams& e, uint32_t top_saddr, int row_shift, uint32_t lut_saddr /*this channel's 256 bytes*/,
                                                 int32_t lo, int sh, const int32_t (&acc)[N], uint32_t (&out)[N]) {
  int32_t a[N];
  uint32_t posb[N];
  const uint32_t base = top_saddr - (1u << row_shift);
  const int levels = e.thr_lut_levels, pmax = e.thr_n + 1 - (1 << levels);
#pragma unroll
  for (int i = 0; i < N; i++) {
    a[i] = wrap_ta(acc[i], e.acc_bits, e.acc_signed);
    const int b = min(max((a[i] - lo) >> sh, 0), 255);
    uint32_t p0;
    asm("ld.shared.u8 %0, [%1];" : "=r"(p0) : "r"(lut_saddr + (uint32_t)b));
    posb[i] = base + ((uint32_t)min((int)p0, pmax) << row_shift);
  }
  const bool strict = (e.cmp == FCB_CMP_LESS) || (e.cmp == FCB_CMP_GREATER_EQUAL);
#pragma unroll 1
  for (int l = 0; l < levels; l++) {
    const uint32_t stepb = ((1u << (levels - 1)) >> l) << row_shift;
    int32_t tv[N];
#pragma unroll
    for (int i = 0; i < N; i++) tv[i] = lds_s32(posb[i] + stepb);
#pragma unroll
    for (int i = 0; i < N; i++)
      if (strict ? (tv[i] < a[i]) : (tv[i] <= a[i])) posb[i] += stepb;
  }
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = thr_finish(e, (int)((posb[i] - base) >> row_shift));
}

// The pooled-threshold instantiations' search (EPI = 1..3 of umma2_conv_kernel): same bucket LUT + LEVELS binary levels as
// activate_thr_lut, with everything a level does not need hoisted out of it.  Preconditions (thrp_extra_warps(), fcb_umma2.cu):
// comp::less / less_equal, 0 <= act_val, act_val + num_th < 256, accumulators already wrapped to TA and within +-2^30.  The
// shared-memory table this runs on was loaded with every threshold decremented when the functor is less_equal (thr <= a  <=>
// thr - 1 < a), so ONE strict compare serves both; INT32_MAX padding never compares below, so the position never passes num_th.
// ROWB (bytes per table row = channels of the CTA x 4) is a compile-time constant: a level is LDS [pos + imm], ISETP, predicated add.
// fin_base = base - act_val * ROWB, so the result is (pos - fin_base) / ROWB with no further add or mask.
__device__ __forceinline__ int32_t min_relu_s32(int32_t a, int32_t b) {  // max(0, min(a, b)): one VIMNMX.RELU
  int32_t d;
  asm("min.s32.relu %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <int N, int LEVELS, int ROWB>
__device__ __forceinline__ void thr_lut_fast(uint32_t base, uint32_t fin_base, uint32_t lut_saddr, int32_t lo, int sh, const int32_t (&a)[N],
                                             uint32_t (&out)[N]) {
  uint32_t posb[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    const int b = min_relu_s32((a[i] - lo) >> sh, 255);
    uint32_t p0;  // sorted index the bucket starts at, pre-clamped to thr_n + 1 - 2^LEVELS at create time
    asm("ld.shared.u8 %0, [%1];" : "=r"(p0) : "r"(lut_saddr + (uint32_t)b));
    posb[i] = base + p0 * (uint32_t)ROWB;
  }
#pragma unroll
  for (int l = 0; l < LEVELS; l++) {
    constexpr uint32_t top = (uint32_t)ROWB << (LEVELS - 1);
    const uint32_t stepb = top >> l;
    int32_t tv[N];
#pragma unroll
    for (int i = 0; i < N; i++) tv[i] = lds_s32(posb[i] + stepb);
#pragma unroll
    for (int i = 0; i < N; i++)
      if (tv[i] < a[i]) posb[i] += stepb;
  }
  constexpr int rsh = ROWB == 512 ? 9 : 10;
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = (posb[i] - fin_base) >> rsh;
}

// Store one output lane per thread of a warp: lane `l` holds channel ch0 + l of one pixel.
// `word` points at that pixel's output word; sub-byte lanes are merged across the warp.
// Must be called by all 32 lanes (uses shuffles); `valid` masks channels >= OFM.
__device__ __forceinline__ void store_lane(uint8_t* word, int ch, bool valid, uint32_t v, int out_bits) {
  const int lane = threadIdx.x & 31;
  if (out_bits == 8) {
    if (valid) word[ch] = (uint8_t)v;
  } else if (out_bits == 16) {
    if (valid) *reinterpret_cast<uint16_t*>(word + 2 * (size_t)ch) = (uint16_t)v;
  } else if (out_bits == 32) {
    if (valid) *reinterpret_cast<uint32_t*>(word + 4 * (size_t)ch) = v;
  } else if (out_bits == 1) {  // one ballot gathers the warp's 32 channel bits; lane 0 stores them
    const uint32_t vm = __ballot_sync(0xffffffffu, valid), bits = __ballot_sync(0xffffffffu, valid && (v & 1u));
    if (lane == 0) {
      uint8_t* dst = word + (ch >> 3);
      if (vm == 0xffffffffu && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) *reinterpret_cast<uint32_t*>(dst) = bits;
      else
        for (int b = 0; b < 4; b++)
          if ((vm >> (8 * b)) & 0xFFu) dst[b] = (uint8_t)(bits >> (8 * b));  // partially valid bytes keep zero pad bits
    }
  } else if (out_bits == 2 || out_bits == 4) {  // 8/out_bits lanes share a byte
    const int per = 8 / out_bits;
    uint32_t b = valid ? (v << ((lane % per) * out_bits)) : 0u;
    for (int s = 1; s < per; s <<= 1) b |= __shfl_xor_sync(0xffffffffu, b, s);
    if (valid && (lane % per) == 0) word[((size_t)ch * out_bits) >> 3] = (uint8_t)b;
  } else {
    // any other lane width (ap_uint<3>, <12>, ...): the warp's 32 lanes form a segment of out_bits 32-bit words that starts on a
    // 4-byte boundary of the stream word (ch - lane is a multiple of 32); lanes may straddle words.  Word j = OR of the pieces of the
    // lanes that touch it (one warp OR-reduction per word), stored by lane j byte by byte up to the last valid channel's bits.
    const uint32_t pos = (uint32_t)lane * (uint32_t)out_bits, w0 = pos >> 5, sh = pos & 31u;
    const uint32_t vv = valid ? (v & ((1u << out_bits) - 1u)) : 0u;
    const uint32_t lo = vv << sh, hi = sh ? (vv >> (32u - sh)) : 0u;
    uint32_t mine = 0;
    for (int j = 0; j < out_bits; j++) {
      const uint32_t c = (w0 == (uint32_t)j ? lo : 0u) | (w0 + 1u == (uint32_t)j ? hi : 0u);
      const uint32_t wj = __reduce_or_sync(0xffffffffu, c);
      if (lane == j) mine = wj;
    }
    const int nvalid = __popc(__ballot_sync(0xffffffffu, valid));   // valid channels are a prefix of the warp
    const int seg_bytes = (nvalid * out_bits + 7) >> 3;
    if (lane < out_bits) {
      uint8_t* dst = word + ((((size_t)(ch - lane)) * out_bits) >> 3) + 4 * lane;
      for (int b = 0; b < 4; b++)
        if (4 * lane + b < seg_bytes) dst[b] = (uint8_t)(mine >> (8 * b));
    }
  }
}

}  // namespace fcb
