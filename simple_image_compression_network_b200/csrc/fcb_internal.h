// fcb_internal.h -- shared declarations of libfinnconv_b200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/finnconv_b200.h"

namespace fcb {

// ---- error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
#define FCB_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      fcb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return FCB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// Experiment switches (environment variables that bend plans for measurements and cross-checks) exist only in experiment builds
// (-DFCB_EXPERIMENT -> tools/libfinnconv_exp.so).  In the product library exp_env() is the constant nullptr: no environment
// variable can change the kernel a caller gets.
#ifdef FCB_EXPERIMENT
inline const char* exp_env(const char* name) { return getenv(name); }
#else
constexpr const char* exp_env(const char*) { return nullptr; }
#endif
inline int exp_int(const char* name, int dflt) {
  const char* v = exp_env(name);
  return v ? atoi(v) : dflt;
}

// ---- derived geometry -------------------------------------------------------------
struct Geom {
  int kind, C, OFM, KX, KY, IX, IY, OX, OY, SX, SY, PAD;
  int DX, DY;                      // dilation (slidingwindow.h:1515-1631), 1 = none
  int pad_l, pad_r, pad_u, pad_d;  // FMPadding_nonsquare split (streamtools.h:374-379); PAD = pad_l when all four are equal
  int engine_hint, pool_signed, pool_min;
  int simd, pe, SF, NF, K;
  int in_bits, in_signed, w_bits, weight_kind;
  int acc_bits, acc_signed, act_kind, out_bits, num_th, act_val, cmp, pool;
  int out_x, out_y;  // after pooling
  size_t in_word_bytes, out_word_bytes, in_img_bytes, out_img_bytes;
  size_t w_word_bytes, weight_bytes, threshold_bytes, bias_bytes;
};
int derive_geom(const fcb_layer_desc* d, Geom* g);  // validates like the reference's CASSERTs
int normalize_desc(const fcb_layer_desc* in, fcb_layer_desc* out);  // either ABI struct size -> the current layout

// ---- epilogue parameters (device-visible POD) -----------------------------------------
struct EpiParams {
  int act_kind, acc_bits, acc_signed, out_bits, num_th, act_val, cmp, pool;
  int ta_bits;  // TA's declared width: accumulators and thresholds lie within it even when the wrap itself was elided (acc_bits = 32)
  const int8_t* bias;    // [OFM]            (FCB_ACT_BIAS_RELU)
  const int32_t* thr;    // [thr_n][thr_stride]: threshold i of channel ch at thr[i*thr_stride + ch]; per channel sorted
                         // ascending, wrapped to TA, padded with INT32_MAX up to thr_n = 2^k - 1 entries (FCB_ACT_THRESHOLDS)
  int thr_n, thr_stride;
  // bucket LUT (optional, NULL if a bucket would hold more than 15 thresholds): for channel ch, bucket b = clamp((a - lo[ch]) >> sh[ch], 0, 255)
  // starts at sorted index lut[ch*256 + b] = #thresholds below the bucket's first value; at most 2^levels - 1 thresholds lie inside it, so
  // `thr_lut_levels` binary levels finish the search
  const uint8_t* thr_lut;
  int thr_lut_levels;  // 3 (<= 7 thresholds per bucket) or 4 (<= 15)
  const int32_t* thr_lo;
  const int32_t* thr_sh;
  const int32_t* thr_cm; // [thr_stride][thr_n + 1] channel-major copy (same values, INT32_MAX padded): the bottom levels of a
                         // search are one aligned 16-byte group of it
};

// ---- engines ----------------------------------------------------------------------
enum Engine { ENG_IMAD = 0, ENG_XNOR = 1, ENG_UMMA = 2, ENG_CHANWISE = 3 };

struct DirectParams {  // imad / xnor_popc direct convolution
  const uint8_t* in;
  uint8_t* out;
  const void* wt;  // imad: int16 [K][OFMp]; dot: dot_pack_weights(); xnor: uint32 [KW][OFMp]
  EpiParams epi;
  int C, OFM, OFMp, KX, KY, DX, DY, IX, IY, OX, OY, SXe, SYe, PAD, PADY, deconv;  // PAD / PADY: zeros left / up of the frame
  int in_bits, in_signed, in_word_bytes, out_word_bytes, out_x, out_y;
  int tiles_x, tiles_y, CC, patch_w, patch_h, mul_kind;
  int dot_pack;  // 0: imad_conv_kernel; 2 / 4: dot_conv_kernel on IDP.2A / IDP.4A (weights fit 8 bits, lanes <= 16 / <= 8 bits)
  unsigned long long in_img_bytes, out_img_bytes;
};
int launch_direct(const DirectParams& p, int engine, int n_images, size_t smem_bytes, cudaStream_t st);
size_t direct_smem_bytes(int engine, int patch_w, int patch_h, int cc);
size_t imad_smem_bytes(int patch_w, int patch_h, int taps, int cc);
int dot_chunk_channels(int patch_w, int patch_h, int taps, int C, int pk, size_t budget);
size_t dot_smem_bytes(int patch_w, int patch_h, int taps, int cc, int pk);
std::vector<uint32_t> dot_pack_weights(const std::vector<int32_t>& W /*[OFM][taps * C]*/, int OFM, int OFMp, int C, int taps, int pk);

// tcgen05 implicit GEMM (fcb_plan.cu: host glue; fcb_umma2.cu: kernels)
struct UmmaPlan;  // opaque to the API file
int umma_eligible(const Geom& g);
int umma_plan_create(const Geom& g, const std::vector<int32_t>& W /*[OFM][K]*/, const EpiParams& epi, int device, UmmaPlan** out);
int umma_plan_create_thin(const Geom& g, const std::vector<int32_t>& W4 /*[OFM][128]*/, const EpiParams& epi, const int8_t* bias_host /*or NULL*/,
                          int device, UmmaPlan** out);
int umma_plan_create_dthin(const Geom& g, const std::vector<int32_t>& W /*[OFM][K]*/, const EpiParams& epi, int device, UmmaPlan** out);
void umma_plan_destroy(UmmaPlan* p);
const char* umma_plan_describe(const UmmaPlan* p);
int umma_run(UmmaPlan* p, const void* d_in, void* d_out, int n_images, cudaStream_t st, uint64_t* launches);

// channel-wise units: depth-wise convolution and Pool_batch (fcb_chanwise.cu)
enum ChanMode { CW_DWCONV = 0, CW_POOL_MAX = 1, CW_POOL_AVG = 2, CW_POOL_ACC = 3, CW_POOL_QUANTAVG = 4 };
struct ChanParams {
  const uint8_t* in;
  uint8_t* out;
  const int16_t* wt;  // depth-wise weights [Kx*Ky][Cpad]
  const uint32_t* wt4; // the same as bytes, four taps per word: [ceil(Kx*Ky / 4)][Cpad] (NULL when a weight does not fit 8 bits)
  EpiParams epi;
  int C, Cpad, KX, KY, DX, DY, IX, IY, OX, OY, SX, SY, pad_l, pad_u;
  int in_bits, in_signed, in_word_bytes, out_word_bytes, out_bits;
  int mode, size, acc_bits, acc_signed;
  int has_init, init;  // CW_POOL_MAX started from `init` instead of the type's minimum (StreamingMaxPool_Precision's min_value)
  unsigned long long in_img_bytes, out_img_bytes;
};
int launch_chanwise(const ChanParams& p, int n_images, cudaStream_t st);
int chanwise_vector_words(const ChanParams& p);  // 4 / 1: byte-lane kernel with 16 / 4 channels per thread; 0: general kernel
int launch_add_streams(const void* d_in1, const void* d_in2, void* d_out, unsigned long long n_words, int ch, int b1, int s1, int b2, int s2, int ob,
                       int offset, int wb1, int wb2, int wbo, cudaStream_t st);

// thin-input lowering (fcb_im2col.cu)
struct Im2colParams {
  const uint8_t* in;
  uint8_t* out;  // [n][OY][OX][128]
  int IX, IY, OX, OY, S, PAD, K, C, KX, in_word_bytes;
  unsigned long long in_img_bytes;
};
int launch_im2col(const Im2colParams& p, int n_images, cudaStream_t st);
// 1-bit lanes -> s8 {-1,+1} NHWC with an explicit -1 border of PAD pixels; out [n][OY][OX][K] (OX = IX+2*PAD, K = padded C)
int launch_expand_bits(const Im2colParams& p, int n_images, cudaStream_t st);

// synthetic data (fcb_synth.cu)
int synth_fill(void* d_ptr, size_t n_bytes, uint64_t seed, uint32_t mask, uint64_t offset, cudaStream_t st);

// Thin-output deconv522, col2im form: index (0..24) of tap t = ky*5 + kx in a pixel's record of 25 four-byte words.  Taps are grouped
// by input shift (offy, offx) = ((k + (k & 1) - 2) / 2): the four 4-tap shifts first (16-byte aligned, words in output-phase order
// ph = 2*(ky&1) + (kx&1)), then the four 2-tap shifts, then tap (0,0).  Weight rows (fcb_umma.cu) and the epilogue (fcb_umma2.cu) share it.
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int dcol_word(int t) {
  const int ky = t / 5, kx = t % 5;
  const int oy = (ky + (ky & 1) - 2) / 2, ox = (kx + (kx & 1) - 2) / 2;
  return (oy >= 0 && ox >= 0) ? (oy * 2 + ox) * 4 + (ky & 1) * 2 + (kx & 1)
         : (oy < 0 && ox >= 0) ? 16 + ox * 2 + (kx & 1)
         : (oy >= 0 && ox < 0) ? 20 + oy * 2 + (ky & 1)
                               : 24;
}

}  // namespace fcb
