/* oracle/finn_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by
 * or called from the product (simple_image_compression_network_b200/ or libfinnconv_b200.so); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * A plain-C CPU restatement of the reference's quantized non-square convolution layer.
 * Parity status: PINNED -- checked against the reference's own templates compiled from
 * /root/reference (oracle/_ref/libref_layers.so, see oracle/Makefile) on every case of
 * oracle/gen_golden.py, and against the frozen golden vectors in tests/golden/.
 *
 * What each step follows:
 *   word images         ap_uint<W> lanes, lane 0 at the LSB   interpret.hpp:191-217, conv3_nonsquare_tb.cpp:807-808
 *   weight tiles        W[nf*PE+pe][sf*SIMD+simd]            mvau.hpp:101-148, weights.hpp:110-150 (binary: :66-98)
 *   padding             zeros, pad/pad split                   streamtools.h:361-406 via conv_nonsquare_top.cpp:59-69
 *   window order        k = (ky*Kx + kx)*C + c                 slidingwindow.h:1302-1313
 *   strided conv        stride-1 windows, keep row%S==col%S==0 conv_nonsquare_top.cpp:238-259
 *   transposed conv     zero-insert, side pad, pad 2, 5x5      conv_nonsquare_top.cpp:109-156
 *   MAC + TA wrap       acc += w*a wrapped to TA               mvau.hpp:122-178, mac.hpp:163-172
 *   xnor / +-1          w==a ? 1:0  /  w ? a : -a              interpret.hpp:57-73, :75-108
 *   pass-through        lane = acc truncated                   activations.hpp:127-134, mvau.hpp:167
 *   bias + ReLU         (lane+bias) mod 2^B, MSB -> 0          conv_nonsquare_top.cpp:267-278
 *   thresholds          ActVal + sum_i cmp(thr_i, acc) in TR   activations.hpp:168-190, :57-99
 *   max pool            k x k stride k, per lane (1-bit: OR)   maxpool.h:137-185, :66-96 (ActType signedness, min_value: :137-170)
 *   padding split       left/up = P/2 (+P%2, style 2)          streamtools.h:374-379
 *   depth-wise conv     window order (chunk, ky, kx); acc[ch] = sum_k W[pe][nf*K2+k] * a_k[ch]   slidingwindow.h:1377-1488, vvau.hpp:80-154
 *   Pool_batch          init / pool / activate per channel    maxpool.h:525-577, pool.hpp:94-226
 * The arithmetic is done in int64 and reduced to TA once before the activation, which is
 * identical to wrapping at every += (two's-complement modular arithmetic).
 */
#include "finn_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void fo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int fo_get_threads(void) { return g_threads; }

size_t fo_word_bytes(uint32_t bits) {
  if (bits <= 8) return 1;
  if (bits <= 16) return 2;
  if (bits <= 32) return 4;
  if (bits <= 64) return 8;
  return 8 * (size_t)((bits + 63) / 64);
}

/* read `n` (<= 32) bits at bit offset `lo` of a little-endian byte image */
static inline uint32_t get_bits(const uint8_t* p, uint64_t lo, uint32_t n) {
  uint64_t v = 0;
  uint64_t byte = lo >> 3;
  uint32_t sh = (uint32_t)(lo & 7);
  uint32_t need = (sh + n + 7) >> 3;
  for (uint32_t i = 0; i < need; i++) v |= (uint64_t)p[byte + i] << (8 * i);
  v >>= sh;
  return (uint32_t)(n >= 32 ? v : (v & ((1ull << n) - 1ull)));
}
static inline void put_bits(uint8_t* p, uint64_t lo, uint32_t n, uint32_t val) {
  for (uint32_t i = 0; i < n; i++) {
    uint64_t b = lo + i;
    uint8_t m = (uint8_t)(1u << (b & 7));
    if ((val >> i) & 1u) p[b >> 3] |= m; else p[b >> 3] &= (uint8_t)~m;
  }
}
static inline int64_t sext(uint64_t v, uint32_t bits) {
  if (bits >= 64) return (int64_t)v;
  uint64_t m = 1ull << (bits - 1);
  v &= (1ull << bits) - 1ull;
  return (int64_t)((v ^ m) - m);
}
static inline int64_t wrap_ta(int64_t v, uint32_t bits, int is_signed) {
  if (bits >= 64) return v;
  uint64_t u = (uint64_t)v & ((1ull << bits) - 1ull);
  return is_signed ? sext(u, bits) : (int64_t)u;
}

/* Dilation_x / Dilation_y (slidingwindow.h:1515-1631: line index (ofm_x*Stride_x + k_x*Dilation_x), block (k_y*Dilation_y)/Stride_y) */
static uint32_t dil_x(const fcb_layer_desc* d) { return d->struct_size == sizeof(fcb_layer_desc) && d->dilation_x > 1 ? d->dilation_x : 1; }
static uint32_t dil_y(const fcb_layer_desc* d) { return d->struct_size == sizeof(fcb_layer_desc) && d->dilation_y > 1 ? d->dilation_y : 1; }

/* FMPadding_nonsquare split (streamtools.h:374-379) */
static void pad_split(const fcb_layer_desc* d, uint32_t* l, uint32_t* r, uint32_t* u, uint32_t* dn) {
  if (!d->pad_style) { *l = *r = *u = *dn = d->pad; return; }
  *l = d->pad_x_total / 2 + (d->pad_style == 2 ? d->pad_x_total % 2 : 0); *r = d->pad_x_total - *l;
  *u = d->pad_y_total / 2 + (d->pad_style == 2 ? d->pad_y_total % 2 : 0); *dn = d->pad_y_total - *u;
}

int fo_layer_query(const fcb_layer_desc* d, fo_sizes* s) {
  if (!d || (d->struct_size != sizeof(fcb_layer_desc) && d->struct_size != FCB_LAYER_DESC_SIZE_V1)) return -1;
  if (!d->simd || !d->pe || !d->ifm_ch || !d->ofm_ch || !d->kernel_x || !d->kernel_y || !d->stride_x || !d->stride_y) return -1;
  if (d->ifm_ch % d->simd) return -2;                                /* slidingwindow.h:1259 */
  if (d->ofm_ch % d->pe) return -2;                                  /* streamtools.h:505 (PE*B -> OFM*B) */
  if (d->in_bits < 1 || d->in_bits > 16 || d->out_bits < 1 || d->out_bits > 32 || d->acc_bits < 1 || d->acc_bits > 64) return -3;
  uint32_t ox, oy, pl, pr, pu, pd;
  pad_split(d, &pl, &pr, &pu, &pd);
  const int chanwise = d->kind == FCB_KIND_DWCONV || d->kind == FCB_KIND_POOL;
  if (d->kind == FCB_KIND_DECONV522) {
    if (d->kernel_x != 5 || d->kernel_y != 5 || d->stride_x != 2 || d->stride_y != 2 || d->pad != 2) return -2;
    ox = 2 * d->ifm_x; oy = 2 * d->ifm_y;
  } else if (d->kind == FCB_KIND_CONV || chanwise) {
    const uint32_t kex = (d->kernel_x - 1) * dil_x(d) + 1, key = (d->kernel_y - 1) * dil_y(d) + 1;
    if (d->ifm_x + pl + pr < kex || d->ifm_y + pu + pd < key) return -2;
    /* kept windows: stride-1 positions 0..I+P-K with pos % S == 0 (conv_nonsquare_top.cpp:246-259) */
    ox = (d->ifm_x + pl + pr - kex) / d->stride_x + 1;
    oy = (d->ifm_y + pu + pd - key) / d->stride_y + 1;
    if (chanwise && (d->ofm_ch != d->ifm_ch || d->simd != d->pe)) return -2;  /* one output channel per input channel; SWG SIMD == PE */
  } else return -1;
  if (ox != d->ofm_x || oy != d->ofm_y) return -2;
  if (d->kind == FCB_KIND_POOL) { if (d->weight_kind > FCB_POOLFN_QUANTAVG) return -1; }
  else if (d->weight_kind == FCB_W_FIXED) { if (d->w_bits < 1 || d->w_bits > 16) return -3; }
  else if (d->weight_kind == FCB_W_BINARY_XNOR) { if (d->w_bits != 1 || d->in_bits != 1) return -2; }
  else if (d->weight_kind == FCB_W_BINARY_PM1) { if (d->w_bits != 1) return -2; }
  else return -1;
  if (d->act_kind > FCB_ACT_THRESHOLDS || d->cmp > FCB_CMP_GREATER_EQUAL) return -1;
  uint32_t pk = d->pool >= 2 ? d->pool : 1;
  if (ox % pk || oy % pk) return -2;                                  /* maxpool.h:140 */
  uint32_t K = d->kernel_x * d->kernel_y * d->ifm_ch;
  if (s) {
    s->k_total = K;
    s->sf = K / d->simd;
    s->nf = d->ofm_ch / d->pe;
    s->out_x = ox / pk; s->out_y = oy / pk;
    s->in_word_bytes = fo_word_bytes(d->ifm_ch * d->in_bits);
    s->out_word_bytes = fo_word_bytes(d->ofm_ch * d->out_bits);
    s->in_bytes_per_image = s->in_word_bytes * d->ifm_x * d->ifm_y;
    s->out_bytes_per_image = s->out_word_bytes * s->out_x * s->out_y;
    s->weight_word_bytes = fo_word_bytes(d->simd * d->w_bits);
    s->weight_bytes = s->weight_word_bytes * d->pe * (size_t)s->sf * s->nf;
    if (d->kind == FCB_KIND_DWCONV) {  /* FixedPointWeights<1, WT, PE, NF*K2>: one lane per word (vvau.hpp:128-134) */
      s->k_total = d->kernel_x * d->kernel_y;
      s->sf = s->k_total;
      s->weight_word_bytes = fo_word_bytes(d->w_bits);
      s->weight_bytes = s->weight_word_bytes * d->pe * (size_t)s->sf * s->nf;
    } else if (d->kind == FCB_KIND_POOL) {
      s->k_total = d->kernel_x * d->kernel_y;
      s->sf = s->k_total;
      s->weight_word_bytes = 0;
      s->weight_bytes = 0;
    }
    s->threshold_bytes = d->act_kind == FCB_ACT_THRESHOLDS ? fo_word_bytes(d->acc_bits) * d->pe * (size_t)s->nf * d->num_th : 0;
    s->bias_bytes = d->act_kind == FCB_ACT_BIAS_RELU ? d->ofm_ch : 0;
  }
  return 0;
}

static int fo_chanwise_run(const fcb_layer_desc* d, const fo_sizes* s, const void* in_words, const void* weights, const void* thresholds,
                           const void* bias, void* out_words, uint32_t numReps);

/* thresholds cmp(thr, acc) -- activations.hpp:57-99,185 */
static inline int cmp_eval(uint32_t cmp, int64_t thr, int64_t acc) {
  switch (cmp) {
    case FCB_CMP_LESS: return thr < acc;
    case FCB_CMP_GREATER: return thr > acc;
    case FCB_CMP_LESS_EQUAL: return thr <= acc;
    default: return thr >= acc;
  }
}

int fo_layer_run(const fcb_layer_desc* d, const void* in_words, const void* weights, const void* thresholds,
                 const void* bias, void* out_words, uint32_t numReps) {
  fo_sizes s;
  int rc = fo_layer_query(d, &s);
  if (rc) return rc;
  if (d->kind == FCB_KIND_DWCONV || d->kind == FCB_KIND_POOL) return fo_chanwise_run(d, &s, in_words, weights, thresholds, bias, out_words, numReps);
  if (!in_words || !weights || !out_words) return -1;
  if (d->act_kind == FCB_ACT_THRESHOLDS && !thresholds) return -1;
  if (d->act_kind == FCB_ACT_BIAS_RELU && !bias) return -1;

  const uint32_t C = d->ifm_ch, OFM = d->ofm_ch, KX = d->kernel_x, KY = d->kernel_y, K = s.k_total;
  const uint32_t OX = d->ofm_x, OY = d->ofm_y;
  const int deconv = d->kind == FCB_KIND_DECONV522;
  /* extent of the fully padded (and, for deconv, zero-inserted) frame the 5x5 / KxK windows slide over */
  uint32_t pl, pr, pu, pd;
  pad_split(d, &pl, &pr, &pu, &pd);
  const uint32_t PX = deconv ? 2 * d->ifm_x + 4 : d->ifm_x + pl + pr;
  const uint32_t PY = deconv ? 2 * d->ifm_y + 4 : d->ifm_y + pu + pd;
  const uint32_t SX = deconv ? 1 : d->stride_x, SY = deconv ? 1 : d->stride_y;
  const uint32_t DXo = deconv ? 1 : dil_x(d), DYo = deconv ? 1 : dil_y(d);

  /* --- weights: W[ch][k] (A.2) ------------------------------------------------ */
  int32_t* W = (int32_t*)malloc(sizeof(int32_t) * (size_t)OFM * K);
  int64_t* TH = NULL;
  int16_t* Wq = NULL;
  if (!W) return -5;
  {
    const uint8_t* wb = (const uint8_t*)weights;
    for (uint32_t pe = 0; pe < d->pe; pe++)
      for (uint32_t nf = 0; nf < s.nf; nf++)
        for (uint32_t sf = 0; sf < s.sf; sf++) {
          const uint8_t* word = wb + ((size_t)pe * s.nf * s.sf + (size_t)nf * s.sf + sf) * s.weight_word_bytes;
          for (uint32_t l = 0; l < d->simd; l++) {
            uint32_t raw = get_bits(word, (uint64_t)l * d->w_bits, d->w_bits);
            int32_t v = d->weight_kind == FCB_W_FIXED ? (int32_t)sext(raw, d->w_bits) : (int32_t)raw;
            W[(size_t)(nf * d->pe + pe) * K + sf * d->simd + l] = v;
          }
        }
  }
  if (d->act_kind == FCB_ACT_THRESHOLDS) {
    TH = (int64_t*)malloc(sizeof(int64_t) * (size_t)OFM * (d->num_th ? d->num_th : 1));
    if (!TH) { free(W); return -5; }
    const uint8_t* tb = (const uint8_t*)thresholds;
    size_t cb = fo_word_bytes(d->acc_bits);
    for (uint32_t pe = 0; pe < d->pe; pe++)
      for (uint32_t nf = 0; nf < s.nf; nf++)
        for (uint32_t i = 0; i < d->num_th; i++) {
          const uint8_t* p = tb + (((size_t)pe * s.nf + nf) * d->num_th + i) * cb;
          uint64_t raw = 0;
          for (size_t b = 0; b < cb && b < 8; b++) raw |= (uint64_t)p[b] << (8 * b);
          TH[(size_t)(nf * d->pe + pe) * d->num_th + i] = wrap_ta((int64_t)raw, d->acc_bits, d->acc_signed);
        }
  }
  /* fast path: 16-bit operands, 32-bit accumulation when the exact sum cannot overflow int32 */
  const int fast = d->weight_kind == FCB_W_FIXED && d->in_bits <= 8 && d->w_bits <= 8 && (uint64_t)K * 256ull * 128ull < (1ull << 31);
  if (fast) {
    Wq = (int16_t*)malloc(sizeof(int16_t) * (size_t)OFM * K);
    if (!Wq) { free(W); free(TH); return -5; }
    for (size_t i = 0; i < (size_t)OFM * K; i++) Wq[i] = (int16_t)W[i];
  }

  const uint32_t pk = d->pool >= 2 ? d->pool : 1;
  const uint64_t out_mask = d->out_bits >= 32 ? 0xffffffffull : ((1ull << d->out_bits) - 1ull);
  int err = 0;

  for (uint32_t n = 0; n < numReps && !err; n++) {
    const uint8_t* img = (const uint8_t*)in_words + (size_t)n * s.in_bytes_per_image;
    uint8_t* oimg = (uint8_t*)out_words + (size_t)n * s.out_bytes_per_image;
    /* padded frame, lanes as int16 (in_bits <= 16 signed or unsigned fits int32; int16 for <= 8/unsigned 15) */
    int32_t* P = (int32_t*)calloc((size_t)PX * PY * C, sizeof(int32_t));
    uint32_t* act = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)OX * OY * OFM); /* pre-pool output lanes */
    if (!P || !act) { free(P); free(act); err = -5; break; }
    for (uint32_t y = 0; y < d->ifm_y; y++)
      for (uint32_t x = 0; x < d->ifm_x; x++) {
        const uint8_t* word = img + ((size_t)y * d->ifm_x + x) * s.in_word_bytes;
        uint32_t px = deconv ? 2 * x + 2 : x + pl; /* A.6: Z(2i,2j) = a(i,j), then pad 2 */
        uint32_t py = deconv ? 2 * y + 2 : y + pu;
        int32_t* dst = P + ((size_t)py * PX + px) * C;
        for (uint32_t c = 0; c < C; c++) {
          uint32_t raw = get_bits(word, (uint64_t)c * d->in_bits, d->in_bits);
          dst[c] = d->in_signed ? (int32_t)sext(raw, d->in_bits) : (int32_t)raw;
        }
      }

#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_threads)
#endif
    for (int oy_ = 0; oy_ < (int)OY; oy_++) {
      uint32_t oy = (uint32_t)oy_;
      int32_t* win = (int32_t*)malloc(sizeof(int32_t) * K);
      int16_t* winq = (int16_t*)malloc(sizeof(int16_t) * K);
      for (uint32_t ox = 0; ox < OX; ox++) {
        /* sliding window in (ky, kx, c) order; dilated taps step DX / DY frame pixels */
        for (uint32_t ky = 0; ky < KY; ky++)
          for (uint32_t kx = 0; kx < KX; kx++) {
            const int32_t* src = P + ((size_t)(oy * SY + ky * DYo) * PX + (size_t)ox * SX + (size_t)kx * DXo) * C;
            memcpy(win + ((size_t)ky * KX + kx) * C, src, sizeof(int32_t) * C);
          }
        if (fast) for (uint32_t k = 0; k < K; k++) winq[k] = (int16_t)win[k];
        uint32_t* o = act + ((size_t)oy * OX + ox) * OFM;
        for (uint32_t ch = 0; ch < OFM; ch++) {
          int64_t acc = 0;
          if (fast) {
            const int16_t* w = Wq + (size_t)ch * K;
            int32_t a32 = 0;
            for (uint32_t k = 0; k < K; k++) a32 += (int32_t)w[k] * (int32_t)winq[k];
            acc = a32;
          } else if (d->weight_kind == FCB_W_FIXED) {
            const int32_t* w = W + (size_t)ch * K;
            for (uint32_t k = 0; k < K; k++) acc += (int64_t)w[k] * (int64_t)win[k];
          } else if (d->weight_kind == FCB_W_BINARY_XNOR) {
            const int32_t* w = W + (size_t)ch * K;
            for (uint32_t k = 0; k < K; k++) acc += (w[k] == win[k]) ? 1 : 0;
          } else {
            const int32_t* w = W + (size_t)ch * K;
            for (uint32_t k = 0; k < K; k++) acc += w[k] ? (int64_t)win[k] : -(int64_t)win[k];
          }
          acc = wrap_ta(acc, d->acc_bits, d->acc_signed);
          uint64_t r;
          if (d->act_kind == FCB_ACT_PASSTHROUGH) {
            r = (uint64_t)acc & out_mask;
          } else if (d->act_kind == FCB_ACT_BIAS_RELU) {
            int64_t b = (int8_t)((const uint8_t*)bias)[ch];
            r = ((uint64_t)acc + (uint64_t)b) & out_mask;
            if ((r >> (d->out_bits - 1)) & 1ull) r = 0;
          } else {
            int64_t cnt = d->act_val;
            const int64_t* t = TH + (size_t)ch * d->num_th;
            for (uint32_t i = 0; i < d->num_th; i++) cnt += cmp_eval(d->cmp, t[i], acc);
            r = (uint64_t)cnt & out_mask;
          }
          o[ch] = (uint32_t)r;
        }
      }
      free(win);
      free(winq);
    }

    /* pool + pack */
    memset(oimg, 0, s.out_bytes_per_image);
    for (uint32_t yp = 0; yp < s.out_y; yp++)
      for (uint32_t xp = 0; xp < s.out_x; xp++) {
        uint8_t* word = oimg + ((size_t)yp * s.out_x + xp) * s.out_word_bytes;
        for (uint32_t ch = 0; ch < OFM; ch++) {
          /* StreamingMaxPool_Precision (maxpool.h:144-170): buf = min_value converted to ActType; `channeldata > oldMax` in ActType
           * (ap_int<out_bits> when pool_signed); 1-bit StreamingMaxPool is an OR (:81-86) == the unsigned max from 0.
           * pk == 1: no pool unit at all. */
          int64_t m = pk > 1 ? wrap_ta((int64_t)d->pool_min_value, d->out_bits, d->pool_signed) : 0;
          int first = pk == 1;
          for (uint32_t ky = 0; ky < pk; ky++)
            for (uint32_t kx = 0; kx < pk; kx++) {
              uint32_t raw = act[((size_t)(yp * pk + ky) * OX + (xp * pk + kx)) * OFM + ch];
              int64_t v = d->pool_signed ? sext(raw, d->out_bits) : (int64_t)raw;
              if (first || v > m) { m = v; first = 0; }
            }
          put_bits(word, (uint64_t)ch * d->out_bits, d->out_bits, (uint32_t)((uint64_t)m & out_mask));
        }
      }
    free(P);
    free(act);
  }
  free(W);
  free(Wq);
  free(TH);
  return err;
}

/* Channel-wise units.  The sliding window is ConvolutionInputGenerator[_NonSquare]_dws (slidingwindow.h:761-868, 1377-1488): per output
 * pixel, per channel chunk, the taps in (ky, kx) order -- so tap k = ky*Kx + kx of channel ch multiplies weight index nf*K2 + k of
 * lane pe (ch = nf*PE + pe) in Vector_Vector_Activate_Batch (vvau.hpp:106-134), or enters function.pool() in Pool_batch
 * (maxpool.h:548-560).  FMPadding_nonsquare zeros in front when padded. */
static int fo_chanwise_run(const fcb_layer_desc* d, const fo_sizes* s, const void* in_words, const void* weights, const void* thresholds,
                           const void* bias, void* out_words, uint32_t numReps) {
  const int pool = d->kind == FCB_KIND_POOL;
  if (!in_words || !out_words || (!pool && !weights)) return -1;
  if (!pool && d->act_kind == FCB_ACT_THRESHOLDS && !thresholds) return -1;
  if (!pool && d->act_kind == FCB_ACT_BIAS_RELU && !bias) return -1;
  const uint32_t C = d->ifm_ch, KX = d->kernel_x, KY = d->kernel_y, K2 = KX * KY, OX = d->ofm_x, OY = d->ofm_y, PE = d->pe;
  uint32_t pl, pr, pu, pd;
  pad_split(d, &pl, &pr, &pu, &pd);
  const uint64_t out_mask = d->out_bits >= 32 ? 0xffffffffull : ((1ull << d->out_bits) - 1ull);
  const size_t tcb = fo_word_bytes(d->acc_bits);
  for (uint32_t n = 0; n < numReps; n++) {
    const uint8_t* img = (const uint8_t*)in_words + (size_t)n * s->in_bytes_per_image;
    uint8_t* oimg = (uint8_t*)out_words + (size_t)n * s->out_bytes_per_image;
    memset(oimg, 0, s->out_bytes_per_image);
    for (uint32_t oy = 0; oy < OY; oy++)
      for (uint32_t ox = 0; ox < OX; ox++) {
        uint8_t* oword = oimg + ((size_t)oy * OX + ox) * s->out_word_bytes;
        for (uint32_t ch = 0; ch < C; ch++) {
          const uint32_t nf = ch / PE, pe = ch % PE;
          int64_t acc;
          if (!pool) acc = 0;                                            /* activation.init: 0 for PassThrough / Thresholds */
          else if (d->weight_kind == FCB_POOLFN_MAX)                     /* pool.hpp:98-102: type minimum */
            acc = d->acc_signed ? -((int64_t)1 << (d->acc_bits - 1)) : 0;
          else acc = 0;                                                  /* pool.hpp:66-69 */
          for (uint32_t ky = 0; ky < KY; ky++)
            for (uint32_t kx = 0; kx < KX; kx++) {
              const int64_t y = (int64_t)oy * d->stride_y + (int64_t)ky * dil_y(d) - pu, x = (int64_t)ox * d->stride_x + (int64_t)kx * dil_x(d) - pl;
              int64_t a = 0;                                             /* FMPadding zero */
              if (y >= 0 && y < (int64_t)d->ifm_y && x >= 0 && x < (int64_t)d->ifm_x) {
                uint32_t raw = get_bits(img + ((size_t)y * d->ifm_x + x) * s->in_word_bytes, (uint64_t)ch * d->in_bits, d->in_bits);
                a = d->in_signed ? sext(raw, d->in_bits) : (int64_t)raw;
              }
              if (!pool) {
                const uint8_t* ww = (const uint8_t*)weights + ((size_t)pe * s->nf * K2 + (size_t)nf * K2 + ky * KX + kx) * s->weight_word_bytes;
                const int64_t w = sext(get_bits(ww, 0, d->w_bits), d->w_bits);
                acc = wrap_ta(acc + w * a, d->acc_bits, d->acc_signed);  /* vvau.hpp:131 */
              } else {
                const int64_t v = wrap_ta(a, d->acc_bits, d->acc_signed); /* the slice converts to the function's type */
                if (d->weight_kind == FCB_POOLFN_MAX) acc = v > acc ? v : acc;         /* pool.hpp:109-112 */
                else acc = wrap_ta(acc + v, d->acc_bits, d->acc_signed);               /* :141-144, :175-178, :211-214 */
              }
            }
          uint64_t r;
          if (pool) {
            int64_t o = acc;
            if (d->weight_kind == FCB_POOLFN_AVG) o = d->act_val ? acc / (int64_t)d->act_val : 0;  /* accu / size, C++ truncation (:151-154) */
            else if (d->weight_kind == FCB_POOLFN_QUANTAVG) o = acc >> d->act_val;                  /* TO(accu >> size) (:221-224) */
            r = (uint64_t)o & out_mask;
          } else if (d->act_kind == FCB_ACT_PASSTHROUGH) {
            r = (uint64_t)acc & out_mask;
          } else if (d->act_kind == FCB_ACT_BIAS_RELU) {
            int64_t b = (int8_t)((const uint8_t*)bias)[ch];
            r = ((uint64_t)acc + (uint64_t)b) & out_mask;
            if ((r >> (d->out_bits - 1)) & 1ull) r = 0;
          } else {
            int64_t cnt = d->act_val;
            for (uint32_t i = 0; i < d->num_th; i++) {
              const uint8_t* p = (const uint8_t*)thresholds + (((size_t)pe * s->nf + nf) * d->num_th + i) * tcb;
              uint64_t raw = 0;
              for (size_t b = 0; b < tcb && b < 8; b++) raw |= (uint64_t)p[b] << (8 * b);
              cnt += cmp_eval(d->cmp, wrap_ta((int64_t)raw, d->acc_bits, d->acc_signed), acc);
            }
            r = (uint64_t)cnt & out_mask;
          }
          put_bits(oword, (uint64_t)ch * d->out_bits, d->out_bits, (uint32_t)r);
        }
      }
  }
  return 0;
}

/* standalone pool on a packed stream (non-square restatement of maxpool.h:137-185 / :66-96) */
int fo_maxpool(const void* in_words, void* out_words, uint32_t dim_x, uint32_t dim_y, uint32_t pool, uint32_t ch, uint32_t bits) {
  if (!pool || dim_x % pool || dim_y % pool) return -2;
  size_t wb = fo_word_bytes(ch * bits);
  uint32_t ox = dim_x / pool, oy = dim_y / pool;
  memset(out_words, 0, wb * ox * oy);
  for (uint32_t yp = 0; yp < oy; yp++)
    for (uint32_t xp = 0; xp < ox; xp++)
      for (uint32_t c = 0; c < ch; c++) {
        uint32_t m = 0;
        for (uint32_t ky = 0; ky < pool; ky++)
          for (uint32_t kx = 0; kx < pool; kx++) {
            const uint8_t* w = (const uint8_t*)in_words + ((size_t)(yp * pool + ky) * dim_x + xp * pool + kx) * wb;
            uint32_t v = get_bits(w, (uint64_t)c * bits, bits);
            if (v > m) m = v;
          }
        put_bits((uint8_t*)out_words + ((size_t)yp * ox + xp) * wb, (uint64_t)c * bits, bits, m);
      }
  return 0;
}

/* AddStreams_Batch (streamtools.h:669-720): per word and channel, Out_t sum = op1 + op2 + offset, operands read as In1_t / In2_t
 * (ap_int or ap_uint of their width), result wrapped to Out_t's width. */
int fo_add_streams(const void* in1, const void* in2, void* out, uint64_t n_words, uint32_t channels, uint32_t in1_bits, int in1_signed,
                   uint32_t in2_bits, int in2_signed, uint32_t out_bits, int32_t offset) {
  if (!in1 || !in2 || !out || !channels || in1_bits < 1 || in1_bits > 32 || in2_bits < 1 || in2_bits > 32 || out_bits < 1 || out_bits > 32) return -1;
  const size_t b1 = fo_word_bytes(channels * in1_bits), b2 = fo_word_bytes(channels * in2_bits), bo = fo_word_bytes(channels * out_bits);
  memset(out, 0, bo * n_words);
  for (uint64_t i = 0; i < n_words; i++)
    for (uint32_t c = 0; c < channels; c++) {
      uint32_t r1 = get_bits((const uint8_t*)in1 + i * b1, (uint64_t)c * in1_bits, in1_bits);
      uint32_t r2 = get_bits((const uint8_t*)in2 + i * b2, (uint64_t)c * in2_bits, in2_bits);
      int64_t a = in1_signed ? sext(r1, in1_bits) : (int64_t)r1, b = in2_signed ? sext(r2, in2_bits) : (int64_t)r2;
      uint64_t sum = (uint64_t)(a + b + (int64_t)offset);
      put_bits((uint8_t*)out + i * bo, (uint64_t)c * out_bits, out_bits, (uint32_t)(out_bits >= 32 ? sum : (sum & ((1ull << out_bits) - 1ull))));
    }
  return 0;
}
