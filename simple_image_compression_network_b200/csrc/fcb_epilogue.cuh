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

__device__ __forceinline__ int32_t wrap_ta(int32_t acc, int bits, int sgn) {
  if (bits >= 32) return acc;
  const uint32_t u = (uint32_t)acc << (32 - bits);
  return sgn ? ((int32_t)u >> (32 - bits)) : (int32_t)(u >> (32 - bits));
}

__device__ __forceinline__ uint32_t activate(const EpiParams& e, int ch, int32_t acc) {
  const int32_t a = wrap_ta(acc, e.acc_bits, e.acc_signed);
  const uint32_t omask = e.out_bits >= 32 ? 0xffffffffu : ((1u << e.out_bits) - 1u);
  if (e.act_kind == FCB_ACT_PASSTHROUGH) return (uint32_t)a & omask;
  if (e.act_kind == FCB_ACT_BIAS_RELU) {
    const uint32_t r = ((uint32_t)a + (uint32_t)(int32_t)e.bias[ch]) & omask;
    return ((r >> (e.out_bits - 1)) & 1u) ? 0u : r;
  }
  const int32_t* __restrict__ t = e.thr + (size_t)ch * e.num_th;
  const bool strict = (e.cmp == FCB_CMP_LESS) || (e.cmp == FCB_CMP_GREATER_EQUAL);
  int lo = 0, hi = e.num_th;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int32_t tv = __ldg(t + mid);
    const bool right = strict ? (tv < a) : (tv <= a);
    if (right) lo = mid + 1; else hi = mid;
  }
  const int cnt = (e.cmp == FCB_CMP_LESS || e.cmp == FCB_CMP_LESS_EQUAL) ? lo : (e.num_th - lo);
  return (uint32_t)(e.act_val + cnt) & omask;
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
  } else {  // 1, 2, 4 bits: 8/out_bits lanes share a byte
    const int per = 8 / out_bits;
    uint32_t b = valid ? (v << ((lane % per) * out_bits)) : 0u;
    for (int s = 1; s < per; s <<= 1) b |= __shfl_xor_sync(0xffffffffu, b, s);
    if (valid && (lane % per) == 0) word[((size_t)ch * out_bits) >> 3] = (uint8_t)b;
  }
}

}  // namespace fcb
