// fcb_api.cu -- the C ABI of include/finnconv_b200.h: descriptor validation, parameter re-layout,
// engine selection, host<->device plumbing.  No arithmetic of the layer happens on the host.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "fcb_internal.h"

namespace fcb {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

static size_t word_bytes(uint32_t bits) {
  if (bits <= 8) return 1;
  if (bits <= 16) return 2;
  if (bits <= 32) return 4;
  if (bits <= 64) return 8;
  return 8 * (size_t)((bits + 63) / 64);
}

// Validation mirrors the reference's preconditions: IFMChannels % SIMD (slidingwindow.h:1259),
// DWC divisibility PE*B | OFM*B (streamtools.h:505), TILES == NF*SF (mvau.hpp:101-105,117), pool
// divisibility (maxpool.h:140), and deconv522's hard-wired k5 s2 p2 (conv_nonsquare_top.cpp:84-86).
// A caller built against ABI 0.1 passes the shorter struct (FCB_LAYER_DESC_SIZE_V1): copy what it has into the current layout, the
// appended fields read as zero.  Nothing past the caller's struct_size is ever touched.
int normalize_desc(const fcb_layer_desc* in, fcb_layer_desc* out) {
  if (!in) { set_error("descriptor is NULL"); return FCB_ERR_INVALID_ARG; }
  if (in->struct_size != sizeof(fcb_layer_desc) && in->struct_size != FCB_LAYER_DESC_SIZE_V1) {
    set_error("struct_size %u is neither %zu nor %zu", in->struct_size, sizeof(fcb_layer_desc), (size_t)FCB_LAYER_DESC_SIZE_V1);
    return FCB_ERR_INVALID_ARG;
  }
  memset(out, 0, sizeof(*out));
  memcpy(out, in, in->struct_size);
  out->struct_size = sizeof(fcb_layer_desc);
  return FCB_OK;
}

int derive_geom(const fcb_layer_desc* d_in, Geom* g) {
  fcb_layer_desc dn;
  int nrc = normalize_desc(d_in, &dn);
  if (nrc) return nrc;
  const fcb_layer_desc* d = &dn;
  const bool has_dil = true;
  const uint32_t DX = has_dil && d->dilation_x > 1 ? d->dilation_x : 1, DY = has_dil && d->dilation_y > 1 ? d->dilation_y : 1;
  if ((DX > 1 || DY > 1) && d->kind == FCB_KIND_DECONV522) { set_error("deconv522 has no dilation"); return FCB_ERR_SHAPE; }
  if (DX > 64 || DY > 64) { set_error("dilation > 64"); return FCB_ERR_UNSUPPORTED; }
  const uint32_t KEX = (d->kernel_x - 1) * DX + 1, KEY = (d->kernel_y - 1) * DY + 1;  // extent of the dilated kernel
  if (d->engine_hint > FCB_ENGINE_TENSOR || d->pad_style > 2) { set_error("bad enum value"); return FCB_ERR_INVALID_ARG; }
  if (d->pad_style && d->pad) { set_error("pad must be 0 when pad_style selects pad_x_total / pad_y_total"); return FCB_ERR_INVALID_ARG; }
  if (!d->pad_style && (d->pad_x_total || d->pad_y_total)) { set_error("pad_x_total / pad_y_total need pad_style 1 or 2"); return FCB_ERR_INVALID_ARG; }
  if (!d->simd || !d->pe || !d->ifm_ch || !d->ofm_ch || !d->kernel_x || !d->kernel_y || !d->stride_x || !d->stride_y || !d->ifm_x || !d->ifm_y) {
    set_error("zero-sized parameter"); return FCB_ERR_INVALID_ARG;
  }
  const bool chanwise = d->kind == FCB_KIND_DWCONV || d->kind == FCB_KIND_POOL;
  if (d->kind > FCB_KIND_POOL || d->act_kind > FCB_ACT_THRESHOLDS || d->cmp > FCB_CMP_GREATER_EQUAL ||
      d->weight_kind > (d->kind == FCB_KIND_POOL ? (uint32_t)FCB_POOLFN_QUANTAVG : (uint32_t)FCB_W_BINARY_PM1)) {
    set_error("bad enum value"); return FCB_ERR_INVALID_ARG;
  }
  if (chanwise) {
    // Vector_Vector_Activate_Batch / Pool_batch work channel by channel: NF = Channels / PE, and the depth-wise generator's SIMD is PE
    if (d->ofm_ch != d->ifm_ch) { set_error("channel-wise unit: ofm_ch %u != ifm_ch %u", d->ofm_ch, d->ifm_ch); return FCB_ERR_SHAPE; }
    if (d->simd != d->pe) { set_error("channel-wise unit: the sliding window's SIMD (%u) must equal PE (%u)", d->simd, d->pe); return FCB_ERR_SHAPE; }
    if (d->pool >= 2) { set_error("channel-wise unit: no fused max pool (chain a FCB_KIND_POOL layer)"); return FCB_ERR_UNSUPPORTED; }
    if (d->kind == FCB_KIND_POOL && d->act_kind != FCB_ACT_PASSTHROUGH) { set_error("FCB_KIND_POOL takes FCB_ACT_PASSTHROUGH"); return FCB_ERR_INVALID_ARG; }
    if (d->kind == FCB_KIND_DWCONV && d->weight_kind != FCB_W_FIXED) { set_error("depth-wise convolution takes FixedPointWeights"); return FCB_ERR_UNSUPPORTED; }
    if (d->kind == FCB_KIND_POOL && d->weight_kind == FCB_POOLFN_QUANTAVG && (d->act_val < 0 || d->act_val > 31)) { set_error("QuantAvgPoolFunction shift out of range"); return FCB_ERR_SHAPE; }
    if (d->kind == FCB_KIND_POOL && d->weight_kind == FCB_POOLFN_AVG && d->act_val <= 0) { set_error("AvgPoolFunction needs size > 0"); return FCB_ERR_SHAPE; }
  }
  if (d->ifm_ch % d->simd) { set_error("IFM_CH %% SIMD != 0 (%u %% %u)", d->ifm_ch, d->simd); return FCB_ERR_SHAPE; }
  if (d->ofm_ch % d->pe) { set_error("OFM_CH %% PE != 0 (%u %% %u)", d->ofm_ch, d->pe); return FCB_ERR_SHAPE; }
  // FMPadding_nonsquare (streamtools.h:374-379): left/up = P/2 (+ P%2 with PaddingStyle 2), right/down = the rest
  uint32_t pl = d->pad, pr = d->pad, pu = d->pad, pd = d->pad;
  if (d->pad_style) {
    pl = d->pad_x_total / 2 + (d->pad_style == 2 ? d->pad_x_total % 2 : 0); pr = d->pad_x_total - pl;
    pu = d->pad_y_total / 2 + (d->pad_style == 2 ? d->pad_y_total % 2 : 0); pd = d->pad_y_total - pu;
  }
  uint32_t ox, oy;
  if (d->kind == FCB_KIND_DECONV522) {
    if (d->kernel_x != 5 || d->kernel_y != 5 || d->stride_x != 2 || d->stride_y != 2 || d->pad != 2) {
      set_error("deconv522 is k5 s2 p2 only"); return FCB_ERR_SHAPE;
    }
    ox = 2 * d->ifm_x; oy = 2 * d->ifm_y;
  } else {
    if (d->ifm_x + pl + pr < KEX || d->ifm_y + pu + pd < KEY) { set_error("kernel larger than padded input"); return FCB_ERR_SHAPE; }
    ox = (d->ifm_x + pl + pr - KEX) / d->stride_x + 1;
    oy = (d->ifm_y + pu + pd - KEY) / d->stride_y + 1;
  }
  if (ox != d->ofm_x || oy != d->ofm_y) { set_error("ofm %ux%u does not match geometry %ux%u", d->ofm_x, d->ofm_y, ox, oy); return FCB_ERR_SHAPE; }
  if (d->kind != FCB_KIND_POOL && d->weight_kind == FCB_W_BINARY_XNOR && (d->w_bits != 1 || d->in_bits != 1)) { set_error("xnor needs 1-bit weights and activations"); return FCB_ERR_SHAPE; }
  if (d->kind != FCB_KIND_POOL && d->weight_kind == FCB_W_BINARY_PM1 && d->w_bits != 1) { set_error("binary weights are 1 bit"); return FCB_ERR_SHAPE; }
  const uint32_t pk = d->pool >= 2 ? d->pool : 1;
  if (ox % pk || oy % pk) { set_error("OFM %% PoolDim != 0"); return FCB_ERR_SHAPE; }
  if (d->act_kind == FCB_ACT_THRESHOLDS && d->num_th == 0) { set_error("thresholds activation with NumTH == 0"); return FCB_ERR_SHAPE; }
  // ---- what this implementation supports (a subset of what the templates can express)
  // lane widths: any ap_uint<N> / ap_int<N> the templates accept (interpret.hpp:191-217) up to 16-bit inputs and 32-bit outputs
  if (d->in_bits < 1 || d->in_bits > 16) { set_error("in_bits %u unsupported (1..16)", d->in_bits); return FCB_ERR_UNSUPPORTED; }
  if (d->out_bits < 1 || d->out_bits > 32) { set_error("out_bits %u unsupported (1..32)", d->out_bits); return FCB_ERR_UNSUPPORTED; }
  if (d->kind != FCB_KIND_POOL && (d->w_bits < 1 || d->w_bits > 16)) { set_error("w_bits %u unsupported (1..16)", d->w_bits); return FCB_ERR_UNSUPPORTED; }
  if (d->in_bits == 16 && !d->in_signed) { /* lanes are staged as int32: fine */ }
  // TA (mvau.hpp:112) up to 64 bits for pass-through and bias+ReLU: their output lanes (<= 32 bits) are the low bits of the exact
  // sum, which 32-bit modular accumulation already carries.  Threshold compares need the whole accumulator: TA <= 32 (31 unsigned).
  if (d->acc_bits < 1 || d->acc_bits > 64 || (d->act_kind == FCB_ACT_THRESHOLDS && (d->acc_bits > 32 || (d->acc_bits == 32 && !d->acc_signed))) ||
      (d->kind == FCB_KIND_POOL && d->acc_bits > 32)) {
    set_error("acc_bits %u unsupported (1..64; thresholds and pool functions: 1..32)", d->acc_bits); return FCB_ERR_UNSUPPORTED;
  }
  if (pk > 16) { set_error("pool %u unsupported (PoolDim <= 16)", d->pool); return FCB_ERR_UNSUPPORTED; }
  if (d->act_kind == FCB_ACT_BIAS_RELU && d->out_bits < 2) { set_error("bias+ReLU needs out_bits >= 2"); return FCB_ERR_UNSUPPORTED; }

  g->kind = d->kind; g->C = d->ifm_ch; g->OFM = d->ofm_ch; g->KX = d->kernel_x; g->KY = d->kernel_y;
  g->IX = d->ifm_x; g->IY = d->ifm_y; g->OX = ox; g->OY = oy; g->SX = d->stride_x; g->SY = d->stride_y; g->PAD = (int)pl;
  g->pad_l = (int)pl; g->pad_r = (int)pr; g->pad_u = (int)pu; g->pad_d = (int)pd;
  g->DX = (int)DX; g->DY = (int)DY;
  g->engine_hint = (int)d->engine_hint; g->pool_signed = d->pool_signed ? 1 : 0; g->pool_min = d->pool_min_value;
  g->simd = d->simd; g->pe = d->pe; g->K = g->KX * g->KY * g->C; g->SF = g->K / g->simd; g->NF = g->OFM / g->pe;
  g->in_bits = d->in_bits; g->in_signed = d->in_signed ? 1 : 0; g->w_bits = d->w_bits; g->weight_kind = d->weight_kind;
  g->acc_bits = d->acc_bits; g->acc_signed = d->acc_signed ? 1 : 0; g->act_kind = d->act_kind; g->out_bits = d->out_bits;
  g->num_th = d->act_kind == FCB_ACT_THRESHOLDS ? d->num_th : 0; g->act_val = d->act_val; g->cmp = d->cmp; g->pool = pk;
  g->out_x = ox / pk; g->out_y = oy / pk;
  g->in_word_bytes = word_bytes(g->C * g->in_bits);
  g->out_word_bytes = word_bytes(g->OFM * g->out_bits);
  g->in_img_bytes = g->in_word_bytes * g->IX * g->IY;
  g->out_img_bytes = g->out_word_bytes * g->out_x * g->out_y;
  g->w_word_bytes = word_bytes(g->simd * g->w_bits);
  g->weight_bytes = g->w_word_bytes * g->pe * (size_t)g->SF * g->NF;
  if (chanwise) {  // Kernel_2 inputs per output lane; FixedPointWeights<1, WT, PE, NF*K2> (vvau.hpp:100-134) or no weights at all (Pool_batch)
    g->K = g->KX * g->KY; g->SF = g->K;
    g->w_word_bytes = d->kind == FCB_KIND_POOL ? 0 : word_bytes(g->w_bits);
    g->weight_bytes = g->w_word_bytes * g->pe * (size_t)g->SF * g->NF;
  }
  g->threshold_bytes = g->act_kind == FCB_ACT_THRESHOLDS ? word_bytes(g->acc_bits) * g->pe * (size_t)g->NF * g->num_th : 0;
  g->bias_bytes = g->act_kind == FCB_ACT_BIAS_RELU ? g->OFM : 0;
  return FCB_OK;
}

static inline uint32_t get_bits(const uint8_t* p, uint64_t lo, uint32_t n) {
  uint64_t v = 0;
  const uint64_t byte = lo >> 3;
  const uint32_t sh = (uint32_t)(lo & 7), need = (sh + n + 7) >> 3;
  for (uint32_t i = 0; i < need; i++) v |= (uint64_t)p[byte + i] << (8 * i);
  v >>= sh;
  return (uint32_t)(n >= 32 ? v : (v & ((1ull << n) - 1ull)));
}
static inline int32_t wrap_host(int64_t v, int bits, int sgn) {
  if (bits >= 32) return (int32_t)v;
  uint32_t u = (uint32_t)v << (32 - bits);
  return sgn ? ((int32_t)u >> (32 - bits)) : (int32_t)(u >> (32 - bits));
}

}  // namespace fcb

using namespace fcb;

namespace {
// Every entry point works on the handle's device and leaves the caller's current device as it found it.
struct DeviceScope {
  int prev = -1;
  cudaError_t err;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    err = (prev == dev) ? cudaSuccess : cudaSetDevice(dev);
  }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
#define FCB_ON_DEVICE(dev)          \
  DeviceScope dev_scope__(dev);     \
  FCB_CUDA_OK(dev_scope__.err)
// images per chunk such that every chunk's byte offset into both streams stays 16-byte aligned (TMA / vector accesses)
size_t align_chunk(size_t chunk, size_t in_img, size_t out_img) {
  size_t m = 1;
  while (((m * in_img) & 15) || ((m * out_img) & 15)) m *= 2;  // m <= 16
  return std::max(m, chunk / m * m);
}
}  // namespace

struct fcb_layer {
  Geom g;   // the layer as described (public sizes)
  Geom gi;  // what the engine computes: == g, or g without its max pool when the pool runs as a separate pass (post_pool)
  bool post_pool = false;
  ChanParams pool_cw{};                       // the pool pass: StreamingMaxPool_Precision as a channel-wise max from min_value
  void* d_unpooled[2] = {nullptr, nullptr};   // the engine's un-pooled output, one buffer per staging slot
  size_t unpooled_imgs = 0;
  fcb_layer_desc desc{};  // as given to fcb_layer_create (fcb_layer_set_params rebuilds from it)
  int device = 0;
  int engine = ENG_IMAD;
  void* d_wt = nullptr;
  void* d_wt4 = nullptr;  // depth-wise weights as bytes, four taps per word (fcb_chanwise.cu)
  int8_t* d_bias = nullptr;
  int32_t* d_thr = nullptr;
  int32_t* d_thr_cm = nullptr;
  uint8_t* d_thr_lut = nullptr;
  int32_t* d_thr_lo = nullptr;  // [2][OFMpad]: lo, then sh
  EpiParams epi{};
  DirectParams dp{};
  ChanParams cw{};
  size_t smem = 0;
  UmmaPlan* umma = nullptr;
  // thin-input layers (Kx*Ky*C <= 128): im2col rows in a scratch buffer, then a 1x1 layer on the tensor-core engine
  bool lowered = false;
  bool lower_bits = false;  // lowering = 1-bit -> +-1 int8 expansion (instead of im2col rows)
  Im2colParams ip{};
  void* d_scratch[2] = {nullptr, nullptr};  // one per staging slot / stream of the host-buffer call (slot 0 serves device calls)
  size_t scratch_imgs = 0, scratch_img_bytes = 0;
  size_t host_chunk = 0;  // fcb_layer_set_host_chunk (0 = default)
  char plan_desc[256] = "";
  // staging for the host-buffer entry point (two slots, double buffered)
  void* s_in[2] = {nullptr, nullptr};
  void* s_out[2] = {nullptr, nullptr};
  size_t s_imgs = 0;
  cudaStream_t s_stream[2] = {nullptr, nullptr};
  uint64_t launches = 0;
};

struct fcb_net {
  std::vector<fcb_layer*> layers;
  std::vector<void*> bufs;  // intermediate activations, one per layer boundary, sized for cap_imgs
  size_t cap_imgs = 0;
  // host-buffer entry point: two staging slots and three streams (H2D | layers | D2H) so the copies of chunk k+1 / k-1 run
  // under the kernels of chunk k
  void* s_in[2] = {nullptr, nullptr};
  void* s_out[2] = {nullptr, nullptr};
  size_t s_imgs = 0;
  cudaStream_t st_in = nullptr, st_run = nullptr, st_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_run[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  int device = 0;
  size_t host_chunk = 0;    // fcb_net_set_host_chunk (0 = default)
  size_t device_chunk = 0;  // fcb_net_set_device_chunk (0 = default)
};

static int layer_run_device(fcb_layer* L, const void* d_in, void* d_out, uint32_t numReps, cudaStream_t st, int slot);
static int lowered_reserve(fcb_layer* L);
static int unpooled_reserve(fcb_layer* L);

extern "C" {

const char* fcb_version(void) { return "finnconv_b200 0.1.0 (sm_100a)"; }
const char* fcb_last_error(void) { return g_err.c_str(); }
size_t fcb_word_bytes(uint32_t bits) { return word_bytes(bits); }

int fcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; i++) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major == 10) ok++;
  }
  return ok;
}

int fcb_layer_query(const fcb_layer_desc* desc, size_t* in_b, size_t* out_b, size_t* w_b, size_t* t_b, size_t* b_b) {
  Geom g;
  int rc = derive_geom(desc, &g);
  if (rc) return rc;
  if (in_b) *in_b = g.in_img_bytes;
  if (out_b) *out_b = g.out_img_bytes;
  if (w_b) *w_b = g.weight_bytes;
  if (t_b) *t_b = g.threshold_bytes;
  if (b_b) *b_b = g.bias_bytes;
  return FCB_OK;
}

void fcb_layer_destroy(fcb_layer* L) {
  if (!L) return;
  DeviceScope ds(L->device);
  if (L->umma) umma_plan_destroy(L->umma);
  cudaFree(L->d_wt); cudaFree(L->d_wt4); cudaFree(L->d_bias); cudaFree(L->d_thr); cudaFree(L->d_thr_cm); cudaFree(L->d_thr_lut); cudaFree(L->d_thr_lo);
  cudaFree(L->d_scratch[0]); cudaFree(L->d_scratch[1]); cudaFree(L->d_unpooled[0]); cudaFree(L->d_unpooled[1]);
  for (int i = 0; i < 2; i++) {
    cudaFree(L->s_in[i]); cudaFree(L->s_out[i]);
    if (L->s_stream[i]) cudaStreamDestroy(L->s_stream[i]);
  }
  delete L;
}

// threshold tables on the device: threshold-major [2^k - 1][stride] for the lock-step search, channel-major copy for the
// 16-byte bottom groups of the hybrid search (fcb_epilogue.cuh)
static int upload_thresholds(fcb_layer* L, const std::vector<std::vector<int32_t>>& rows) {
  const int ofm = (int)rows.size(), nth = ofm ? (int)rows[0].size() : 0;
  int tn = 1;
  while (tn - 1 < nth) tn *= 2;
  tn -= 1;
  const int tstride = (ofm + 127) / 128 * 128;
  std::vector<int32_t> T((size_t)tn * tstride, 0x7fffffff), TC((size_t)tstride * (tn + 1), 0x7fffffff);
  for (int ch = 0; ch < ofm; ch++)
    for (int i = 0; i < nth; i++) {
      T[(size_t)i * tstride + ch] = rows[ch][i];
      TC[(size_t)ch * (tn + 1) + i] = rows[ch][i];
    }
  cudaFree(L->d_thr); cudaFree(L->d_thr_cm);
  L->d_thr = L->d_thr_cm = nullptr;
  FCB_CUDA_OK(cudaMalloc(&L->d_thr, T.size() * sizeof(int32_t)));
  FCB_CUDA_OK(cudaMemcpy(L->d_thr, T.data(), T.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  FCB_CUDA_OK(cudaMalloc(&L->d_thr_cm, TC.size() * sizeof(int32_t)));
  FCB_CUDA_OK(cudaMemcpy(L->d_thr_cm, TC.data(), TC.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  L->epi.thr = L->d_thr; L->epi.thr_cm = L->d_thr_cm; L->epi.thr_n = tn; L->epi.thr_stride = tstride;
  // bucket LUT for the 255-threshold class of tables (see EpiParams): 256 buckets of width 2^sh from the smallest threshold
  cudaFree(L->d_thr_lut); cudaFree(L->d_thr_lo);
  L->d_thr_lut = nullptr; L->d_thr_lo = nullptr;
  L->epi.thr_lut = nullptr; L->epi.thr_lo = L->epi.thr_sh = nullptr;
  // (a - lo) is formed in 32 bits on the device: accumulators and thresholds must stay within +-2^30 (TA of at most 31 bits)
  if (nth >= 64 && nth <= 255 && L->g.acc_bits <= 30) {
    std::vector<uint8_t> lut((size_t)tstride * 256, 0);
    std::vector<int32_t> losh((size_t)2 * tstride, 0);
    bool ok = true;
    int worst = 0;
    for (int ch = 0; ch < ofm && ok; ch++) {
      const std::vector<int32_t>& r = rows[ch];
      const int64_t lo = r.front(), span = (int64_t)r.back() - lo;
      if (lo < -(1ll << 30) || r.back() > (1ll << 30)) { ok = false; break; }
      int sh = 0;
      while ((span >> sh) > 255) sh++;
      losh[ch] = (int32_t)lo; losh[tstride + ch] = sh;
      int idx = 0;
      for (int b = 0; b < 256; b++) {
        const int64_t start = lo + ((int64_t)b << sh);
        while (idx < nth && r[idx] < start) idx++;
        lut[(size_t)ch * 256 + b] = (uint8_t)idx;
        int64_t end = lo + ((int64_t)(b + 1) << sh);
        int inside = 0;
        for (int j = idx; j < nth && (b == 255 || r[j] < end); j++) inside++;
        worst = std::max(worst, inside);
        if (inside > 15) { ok = false; break; }
      }
    }
    if (ok) {
      // entries are clamped so that the 2^levels-wide search window stays inside the 2^D - 1 table rows (thr_lut_fast relies on it;
      // activate_thr_lut clamps again).  Exact: thresholds below the clamped index still all compare below.
      const int levels = worst <= 7 ? 3 : 4, pmax = tn + 1 - (1 << levels);
      for (auto& v : lut) v = (uint8_t)std::min<int>(v, pmax);
      FCB_CUDA_OK(cudaMalloc(&L->d_thr_lut, lut.size()));
      FCB_CUDA_OK(cudaMemcpy(L->d_thr_lut, lut.data(), lut.size(), cudaMemcpyHostToDevice));
      FCB_CUDA_OK(cudaMalloc(&L->d_thr_lo, losh.size() * sizeof(int32_t)));
      FCB_CUDA_OK(cudaMemcpy(L->d_thr_lo, losh.data(), losh.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      L->epi.thr_lut_levels = worst <= 7 ? 3 : 4;
      L->epi.thr_lut = L->d_thr_lut; L->epi.thr_lo = L->d_thr_lo; L->epi.thr_sh = L->d_thr_lo + tstride;
    }
  }
  return FCB_OK;
}

int fcb_layer_create(const fcb_layer_desc* desc, const void* weights, const void* thresholds, const void* bias, int device, fcb_layer** out) {
  if (!out) { set_error("out is NULL"); return FCB_ERR_INVALID_ARG; }
  *out = nullptr;
  Geom g;
  int rc = derive_geom(desc, &g);
  if (rc) return rc;
  if (!weights && g.kind != FCB_KIND_POOL) { set_error("weights is NULL"); return FCB_ERR_INVALID_ARG; }
  if (g.act_kind == FCB_ACT_THRESHOLDS && !thresholds) { set_error("thresholds is NULL"); return FCB_ERR_INVALID_ARG; }
  if (g.act_kind == FCB_ACT_BIAS_RELU && !bias) { set_error("bias is NULL"); return FCB_ERR_INVALID_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("no usable CUDA device %d (this library has no CPU path)", device);
    return FCB_ERR_CUDA;
  }
  cudaDeviceProp prop;
  FCB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return FCB_ERR_CUDA; }
  FCB_ON_DEVICE(device);

  fcb_layer* L = new fcb_layer();
  // every early return below (FCB_CUDA_OK, explicit error returns) releases the half-built layer; `armed` is cleared on success.
  // (paths that call fcb_layer_destroy(L) themselves disarm first)
  struct Guard {
    fcb_layer*& L;
    bool armed = true;
    ~Guard() { if (armed && L) fcb_layer_destroy(L); }
  } guard{L};
  L->g = g;
  normalize_desc(desc, &L->desc);
  L->device = device;
  // StreamingMaxPool[_Precision] (maxpool.h:66-185) is fused into the epilogues for the common form: 2x2 (tensor engine) / 2x2 and 4x4
  // (direct engines), unsigned lanes, maxima starting from 0, behind a conv2d.  Every other form the template expresses -- other
  // PoolDim, signed ActType, a min_value that does not wrap to 0, a pool behind deconv522 -- runs the layer un-pooled into a scratch
  // stream and then the pool as its own streaming pass (fcb_chanwise.cu).
  {
    const uint32_t omask = g.out_bits >= 32 ? 0xffffffffu : ((1u << g.out_bits) - 1u);
    const bool min_is_zero = g.pool_signed ? false : (((uint32_t)g.pool_min) & omask) == 0;
    if (g.pool >= 2 && (g.kind == FCB_KIND_DECONV522 || !(g.pool == 2 || g.pool == 4) || g.pool_signed || !min_is_zero)) {
      L->post_pool = true;
      ChanParams& c = L->pool_cw;
      c.C = g.OFM; c.Cpad = (g.OFM + 31) / 32 * 32; c.KX = c.KY = g.pool; c.SX = c.SY = g.pool; c.DX = c.DY = 1; c.IX = g.OX; c.IY = g.OY; c.OX = g.out_x; c.OY = g.out_y;
      c.pad_l = c.pad_u = 0; c.in_bits = g.out_bits; c.in_signed = g.pool_signed; c.in_word_bytes = (int)g.out_word_bytes;
      c.out_word_bytes = (int)g.out_word_bytes; c.out_bits = g.out_bits; c.acc_bits = g.out_bits; c.acc_signed = g.pool_signed;
      c.mode = CW_POOL_MAX; c.has_init = 1; c.init = wrap_host(g.pool_min, g.out_bits, g.pool_signed);
      c.in_img_bytes = g.out_word_bytes * (size_t)g.OX * g.OY; c.out_img_bytes = g.out_img_bytes;
      g.pool = 1; g.out_x = g.OX; g.out_y = g.OY; g.out_img_bytes = c.in_img_bytes;  // the engine below sees the layer without its pool
    }
  }
  L->gi = g;

  // ---- weights: m_weights[pe][nf*SF+sf] lanes -> W[ch][k] (mvau.hpp:117,148; weights.hpp:134-140)
  std::vector<int32_t> W((size_t)g.OFM * g.K);
  if (g.kind == FCB_KIND_DWCONV) {  // one lane per word: W[ch = nf*PE + pe][k = ky*Kx + kx] = m_weights[pe][nf*K2 + k] (vvau.hpp:106-134)
    const uint8_t* wb = (const uint8_t*)weights;
    for (int pe = 0; pe < g.pe; pe++)
      for (int t = 0; t < g.NF * g.K; t++)
        W[(size_t)((t / g.K) * g.pe + pe) * g.K + t % g.K] =
            wrap_host(get_bits(wb + ((size_t)pe * g.NF * g.K + t) * g.w_word_bytes, 0, g.w_bits), g.w_bits, 1);
  } else if (g.kind != FCB_KIND_POOL) {
    const uint8_t* wb = (const uint8_t*)weights;
    for (int pe = 0; pe < g.pe; pe++)
      for (int nf = 0; nf < g.NF; nf++)
        for (int sf = 0; sf < g.SF; sf++) {
          const uint8_t* word = wb + ((size_t)pe * g.NF * g.SF + (size_t)nf * g.SF + sf) * g.w_word_bytes;
          for (int l = 0; l < g.simd; l++) {
            const uint32_t raw = get_bits(word, (uint64_t)l * g.w_bits, g.w_bits);
            int32_t v;
            if (g.weight_kind == FCB_W_FIXED) v = wrap_host(raw, g.w_bits, 1);
            else if (g.weight_kind == FCB_W_BINARY_PM1) v = raw ? 1 : -1;  // interpret.hpp:87-90
            else v = (int32_t)raw;
            W[(size_t)(nf * g.pe + pe) * g.K + sf * g.simd + l] = v;
          }
        }
  }
  // ---- activation parameters
  L->epi.ta_bits = g.acc_bits;
  L->epi.act_kind = g.act_kind; L->epi.acc_bits = g.acc_bits; L->epi.acc_signed = g.acc_signed; L->epi.out_bits = g.out_bits;
  L->epi.num_th = g.num_th; L->epi.act_val = g.act_val; L->epi.cmp = g.cmp; L->epi.pool = g.pool;
  if (g.act_kind == FCB_ACT_BIAS_RELU) {
    FCB_CUDA_OK(cudaMalloc(&L->d_bias, g.OFM));
    FCB_CUDA_OK(cudaMemcpy(L->d_bias, bias, g.OFM, cudaMemcpyHostToDevice));
    L->epi.bias = L->d_bias;
  }
  // The TA wrap (mvau.hpp:112: every += wraps to TA) is the identity when no partial sum can leave TA's range: with
  // |acc| <= max_ch sum_k |w| * max|a| < 2^(acc_bits-1) the device epilogue may treat the accumulator as a plain int32.
  if (g.acc_signed && g.acc_bits < 32 && g.weight_kind == FCB_W_FIXED && g.kind != FCB_KIND_POOL) {
    const uint64_t amax = g.in_signed ? (1ull << (g.in_bits - 1)) : ((1ull << g.in_bits) - 1);
    uint64_t worst = 0;
    for (int ch = 0; ch < g.OFM; ch++) {
      uint64_t sum = 0;
      for (int k = 0; k < g.K; k++) sum += (uint64_t)std::abs(W[(size_t)ch * g.K + k]);
      worst = std::max(worst, sum);
    }
    if (worst * amax < (1ull << (g.acc_bits - 1))) L->epi.acc_bits = 32;
  }
  // thresholds: wrapped to TA and sorted per channel (the reference result is a count, activations.hpp:181-189)
  std::vector<std::vector<int32_t>> thr_rows;
  if (g.act_kind == FCB_ACT_THRESHOLDS) {
    thr_rows.assign(g.OFM, std::vector<int32_t>(g.num_th));
    const uint8_t* tb = (const uint8_t*)thresholds;
    const size_t cb = word_bytes(g.acc_bits);
    for (int pe = 0; pe < g.pe; pe++)
      for (int nf = 0; nf < g.NF; nf++) {
        std::vector<int32_t>& row = thr_rows[nf * g.pe + pe];
        for (int i = 0; i < g.num_th; i++) {
          const uint8_t* p = tb + (((size_t)pe * g.NF + nf) * g.num_th + i) * cb;
          uint64_t raw = 0;
          for (size_t b = 0; b < cb && b < 8; b++) raw |= (uint64_t)p[b] << (8 * b);
          row[i] = wrap_host((int64_t)raw, g.acc_bits, g.acc_signed);
        }
        std::sort(row.begin(), row.end());
      }
    rc = upload_thresholds(L, thr_rows);
    if (rc) return rc;
  }

  if (g.kind == FCB_KIND_DWCONV || g.kind == FCB_KIND_POOL) {
    // channel-wise units (fcb_chanwise.cu): streaming kernels, one engine
    if (g.engine_hint == FCB_ENGINE_XNOR_POPC || g.engine_hint == FCB_ENGINE_TENSOR) { set_error("channel-wise units run on the CUDA cores only"); return FCB_ERR_UNSUPPORTED; }
    ChanParams& c = L->cw;
    c.C = g.C; c.Cpad = (g.C + 31) / 32 * 32; c.KX = g.KX; c.KY = g.KY; c.IX = g.IX; c.IY = g.IY; c.OX = g.OX; c.OY = g.OY; c.SX = g.SX; c.SY = g.SY;
    c.DX = g.DX; c.DY = g.DY;
    c.pad_l = g.pad_l; c.pad_u = g.pad_u; c.in_bits = g.in_bits; c.in_signed = g.in_signed; c.in_word_bytes = (int)g.in_word_bytes;
    c.out_word_bytes = (int)g.out_word_bytes; c.out_bits = g.out_bits; c.acc_bits = g.acc_bits; c.acc_signed = g.acc_signed;
    c.in_img_bytes = g.in_img_bytes; c.out_img_bytes = g.out_img_bytes; c.size = g.act_val;
    c.mode = g.kind == FCB_KIND_DWCONV ? CW_DWCONV : CW_POOL_MAX + g.weight_kind;
    if (g.kind == FCB_KIND_DWCONV) {
      std::vector<int16_t> Wt((size_t)g.K * c.Cpad, 0);
      for (int ch = 0; ch < g.C; ch++)
        for (int k = 0; k < g.K; k++) Wt[(size_t)k * c.Cpad + ch] = (int16_t)W[(size_t)ch * g.K + k];
      FCB_CUDA_OK(cudaMalloc(&L->d_wt, Wt.size() * 2));
      FCB_CUDA_OK(cudaMemcpy(L->d_wt, Wt.data(), Wt.size() * 2, cudaMemcpyHostToDevice));
      c.wt = (const int16_t*)L->d_wt;
      // byte form for the dot-product inner loop of chanwise_bytes_kernel: word [chunk][ch] = weights of taps 4*chunk .. 4*chunk+3
      bool w8 = true;
      for (int32_t w : W) if (w < -128 || w > 127) { w8 = false; break; }
      if (w8) {
        const int chunks = (g.K + 3) / 4;
        std::vector<uint32_t> W4((size_t)chunks * c.Cpad, 0u);
        for (int ch = 0; ch < g.C; ch++)
          for (int k = 0; k < g.K; k++) W4[(size_t)(k >> 2) * c.Cpad + ch] |= (uint32_t)(uint8_t)(int8_t)W[(size_t)ch * g.K + k] << (8 * (k & 3));
        FCB_CUDA_OK(cudaMalloc(&L->d_wt4, W4.size() * 4));
        FCB_CUDA_OK(cudaMemcpy(L->d_wt4, W4.data(), W4.size() * 4, cudaMemcpyHostToDevice));
        c.wt4 = (const uint32_t*)L->d_wt4;
      }
    }
    c.epi = L->epi;
    L->engine = ENG_CHANWISE;
    if (L->post_pool && (rc = unpooled_reserve(L))) return rc;
    guard.armed = false;
    *out = L;
    return FCB_OK;
  }

  // ---- engine selection (fcb_engine_hint: the reference's resource argument R, mvau.hpp:87-98 -- never changes the result)
  const int hint = g.engine_hint;
  auto env_is = [](const char* name, const char* want) { const char* v = exp_env(name); return v && !strcmp(v, want); };
  const bool force_imad = hint == FCB_ENGINE_IMAD || env_is("FCB_FORCE_ENGINE", "imad");
  int engine = ENG_IMAD;
  const bool dense_bits = (g.in_word_bytes * 8 == (size_t)g.C * g.in_bits);
  const bool xnor_ok = g.weight_kind == FCB_W_BINARY_XNOR && g.C % 32 == 0 && dense_bits;
  if (xnor_ok) engine = ENG_XNOR;
  if (umma_eligible(g)) engine = ENG_UMMA;
  if (force_imad) engine = ENG_IMAD;
  if (hint == FCB_ENGINE_XNOR_POPC) {
    if (!xnor_ok) { set_error("engine_hint XNOR_POPC needs FCB_W_BINARY_XNOR with IFM_CH %% 32 == 0"); return FCB_ERR_UNSUPPORTED; }
    engine = ENG_XNOR;
  }
  L->engine = engine;

  // FCB_ENGINE_TENSOR on a 1-bit xnor layer: the layer on the tensor cores.  With a^ = 2a-1, w^ = 2w-1 in {-1,+1}:
  // sum_k [w_k == a_k] = (K + sum_k a^_k w^_k) / 2 (interpret.hpp:57-73), so thr < matches  <=>  2*thr - K < sum a^w^ :
  // bits are expanded to s8 (a zero-padded border bit is an ordinary 0 activation = -1), thresholds are remapped once,
  // and the layer runs on umma_i8.
  // FCB_ENGINE_AUTO takes this form for the shape class it was measured on -- one threshold, 1-bit lanes out, >= 32 channels in
  // (config 3: 566 k img/s against 182 k on the popcount engine, profiles/r02_xnor_tensor.log) -- unless the experiment build says
  // otherwise; FCB_ENGINE_XNOR_POPC keeps the XNOR/popc warp kernels north_star names.
  const bool auto_xnor_tensor = hint == FCB_ENGINE_AUTO && L->epi.thr_n == 1 && g.out_bits == 1 && g.C >= 32 && !force_imad &&
                                !env_is("FCB_XNOR_ENGINE", "popc");
  const bool want_xnor_tensor = g.weight_kind == FCB_W_BINARY_XNOR && (hint == FCB_ENGINE_TENSOR || env_is("FCB_XNOR_ENGINE", "tensor") || auto_xnor_tensor);
  bool xnor_tensor_done = false;
  if (want_xnor_tensor && g.act_kind == FCB_ACT_THRESHOLDS && g.kind == FCB_KIND_CONV && dense_bits && g.SX == g.SY && g.SX == 1 && g.OFM <= 256 &&
      g.pool <= 2 && g.pad_l == g.pad_r && g.pad_u == g.pad_d && g.pad_l == g.pad_u && g.DX == 1 && g.DY == 1 &&
      (g.acc_bits >= 31 || g.K < (1 << (g.acc_bits - (g.acc_signed ? 1 : 0))))) {
    const int Cp = (g.C + 15) / 16 * 16;
    // Pixel pairs (experiment build, FCB_XNOR_PAIR=1): with <= 64 channels a 128-byte K-block of the tensor kernel is half zeros.  The
    // expansion can write, for frame pixel x, the bytes of pixels x and x+1 back to back (one "pair row"); the layer then has 2*Cp
    // channels, ceil(KX / 2) horizontal taps and dilation 2 (tap kx' = kernel columns 2kx', 2kx'+1; zero weights past KX; odd KX keeps
    // the output width): 6 K-blocks instead of 9 for a 3x3 kernel.  Bit-exact, but measured SLOWER on config 3 (440 k vs 540 k img/s,
    // profiles/r02_xnor_tensor.log): the kernel is not MMA-bound there and the pair rows double the scratch and plane traffic.
    const int pair = (Cp <= 64 && g.KX > 1 && (g.KX & 1) && exp_int("FCB_XNOR_PAIR", 0)) ? 2 : 1;
    const int C2 = pair * Cp, KX2 = (g.KX + pair - 1) / pair;
    Geom g2 = g;
    g2.C = C2; g2.KX = KX2; g2.DX = pair; g2.K = KX2 * g.KY * C2; g2.IX = g.IX + 2 * g.PAD; g2.IY = g.IY + 2 * g.PAD; g2.PAD = 0;
    g2.pad_l = g2.pad_r = g2.pad_u = g2.pad_d = 0;
    g2.in_bits = 8; g2.in_signed = 1; g2.w_bits = 8; g2.weight_kind = FCB_W_FIXED; g2.acc_bits = 32; g2.acc_signed = 1;
    g2.in_word_bytes = C2; g2.in_img_bytes = (size_t)C2 * g2.IX * g2.IY;
    std::vector<int32_t> W2((size_t)g.OFM * g2.K, 0);
    for (int ch = 0; ch < g.OFM; ch++)
      for (int ky = 0; ky < g.KY; ky++)
        for (int kx = 0; kx < g.KX; kx++)
          for (int c = 0; c < g.C; c++)
            W2[(size_t)ch * g2.K + (size_t)(ky * KX2 + kx / pair) * C2 + (kx % pair) * Cp + c] =
                W[(size_t)ch * g.K + (size_t)(ky * g.KX + kx) * g.C + c] ? 1 : -1;
    // t' = 2t - K in 64 bits: the remapped table must stay inside int32 (and below the INT32_MAX padding value)
    std::vector<std::vector<int32_t>> rows2 = thr_rows;
    bool remap_ok = true;
    for (auto& r : rows2)
      for (auto& t : r) {
        const int64_t t2 = 2 * (int64_t)t - g.K;
        if (t2 <= INT32_MIN || t2 >= INT32_MAX) remap_ok = false;
        t = (int32_t)t2;
      }
    if (remap_ok && umma_eligible(g2)) {
      rc = upload_thresholds(L, rows2);
      if (rc) return rc;
      EpiParams e2 = L->epi;
      e2.acc_bits = 32; e2.acc_signed = 1; e2.ta_bits = 32;
      rc = umma_plan_create(g2, W2, e2, device, &L->umma);
      if (rc == FCB_OK) {
        L->epi = e2;
        engine = L->engine = ENG_UMMA;
        L->lowered = true;
        L->lower_bits = true;
        L->scratch_img_bytes = g2.in_img_bytes;
        Im2colParams& ip = L->ip;
        ip.IX = g.IX; ip.IY = g.IY; ip.OX = g2.IX; ip.OY = g2.IY; ip.S = pair; ip.PAD = g.PAD; ip.K = Cp; ip.C = g.C; ip.KX = g.KX;  // (S: pixels per row)
        ip.in_word_bytes = (int)g.in_word_bytes; ip.in_img_bytes = g.in_img_bytes;
        snprintf(L->plan_desc, sizeof(L->plan_desc), "xnor as +-1 int8: bit expansion (C=%d -> %d B%s) + %s", g.C, Cp,
                 pair == 2 ? ", pixel pairs per row" : "", umma_plan_describe(L->umma));
        xnor_tensor_done = true;
      } else {
        L->umma = nullptr;
        int rc3 = upload_thresholds(L, thr_rows);  // back to the popcount engine's tables
        if (rc3) return rc3;
        if (rc != FCB_ERR_UNSUPPORTED) return rc;
      }
    }
  }
  if (hint == FCB_ENGINE_TENSOR && g.weight_kind == FCB_W_BINARY_XNOR && !xnor_tensor_done) {
    set_error("engine_hint TENSOR: this xnor layer has no tensor-core form (needs thresholds, stride 1, OFM <= 256, symmetric padding)");
    return FCB_ERR_UNSUPPORTED;
  }
  if (engine == ENG_UMMA && !L->lowered && g.kind == FCB_KIND_DECONV522 && g.OFM <= 4) {
    // thin-output transposed conv (the 3-channel last layer): dedicated plan, pixels on the MMA M axis
    rc = umma_plan_create_dthin(g, W, L->epi, device, &L->umma);
    if (rc == FCB_ERR_UNSUPPORTED) L->umma = nullptr;
    else if (rc) return rc;
  }
  if (engine == ENG_UMMA && !L->lowered && !L->umma) {
    rc = umma_plan_create(g, W, L->epi, device, &L->umma);
    if (rc == FCB_ERR_UNSUPPORTED) { engine = L->engine = ENG_IMAD; L->umma = nullptr; }  // shape the planner cannot tile
    else if (rc) return rc;
  }
  // thin-input layers (one 4-byte word per pixel, e.g. the ap_uint<24> C = 3 first layer): the sliding window is built in
  // shared memory inside the tensor-core kernel (fcb_umma2.cu, thin-input mode); FCB_THIN=im2col keeps the two-kernel lowering
  const bool sym_pad = g.pad_l == g.pad_r && g.pad_u == g.pad_d && g.pad_l == g.pad_u && g.DX == 1 && g.DY == 1;  // (and no dilation)
  if (engine == ENG_IMAD && !force_imad && !env_is("FCB_THIN", "im2col") && g.kind == FCB_KIND_CONV && sym_pad &&
      g.weight_kind == FCB_W_FIXED && g.w_bits <= 8 && g.in_bits == 8 && g.in_word_bytes == 4 && g.KX * g.KY <= 32 && g.OFM <= 256 &&
      g.pool <= 2 && g.SX == g.SY && g.SX <= 2 && g.IX % 4 == 0 && (uint64_t)g.K * 255ull * 128ull < (1ull << 31)) {
    std::vector<int32_t> W4((size_t)g.OFM * 128, 0);
    for (int ch = 0; ch < g.OFM; ch++)
      for (int tap = 0; tap < g.KX * g.KY; tap++)
        for (int c = 0; c < g.C; c++) W4[(size_t)ch * 128 + tap * 4 + c] = W[(size_t)ch * g.K + tap * g.C + c];
    rc = umma_plan_create_thin(g, W4, L->epi, g.act_kind == FCB_ACT_BIAS_RELU ? (const int8_t*)bias : nullptr, device, &L->umma);
    if (rc == FCB_OK) engine = L->engine = ENG_UMMA;
    else if (rc != FCB_ERR_UNSUPPORTED) return rc;
    else L->umma = nullptr;
  }
  // older two-kernel lowering: conv2d with Kx*Ky*C <= 128 bytes of window as im2col rows + a 1x1 layer
  if (engine == ENG_IMAD && !force_imad && g.kind == FCB_KIND_CONV && g.weight_kind == FCB_W_FIXED && sym_pad &&
      g.w_bits <= 8 && g.in_bits == 8 && g.K <= 128 && g.OFM <= 256 && g.pool <= 2 && g.SX == g.SY) {
    Geom g2 = g;
    g2.C = 128; g2.KX = g2.KY = 1; g2.K = 128; g2.IX = g.OX; g2.IY = g.OY; g2.SX = g2.SY = 1; g2.PAD = 0; g2.pad_l = g2.pad_r = g2.pad_u = g2.pad_d = 0;
    g2.in_word_bytes = 128; g2.in_img_bytes = (size_t)128 * g.OX * g.OY;
    std::vector<int32_t> W2((size_t)g.OFM * 128, 0);
    for (int ch = 0; ch < g.OFM; ch++)
      for (int k = 0; k < g.K; k++) W2[(size_t)ch * 128 + k] = W[(size_t)ch * g.K + k];
    if (umma_eligible(g2)) {
      rc = umma_plan_create(g2, W2, L->epi, device, &L->umma);
      if (rc == FCB_OK) {
        engine = L->engine = ENG_UMMA;
        L->lowered = true;
        L->scratch_img_bytes = g2.in_img_bytes;
        Im2colParams& ip = L->ip;
        ip.IX = g.IX; ip.IY = g.IY; ip.OX = g.OX; ip.OY = g.OY; ip.S = g.SX; ip.PAD = g.PAD; ip.K = g.K; ip.C = g.C; ip.KX = g.KX;
        ip.in_word_bytes = (int)g.in_word_bytes; ip.in_img_bytes = g.in_img_bytes;
        snprintf(L->plan_desc, sizeof(L->plan_desc), "im2col rows (K=%d -> 128 B) + 1x1 %s", g.K, umma_plan_describe(L->umma));
      } else if (rc != FCB_ERR_UNSUPPORTED) return rc;
      else L->umma = nullptr;
    }
  }
  if (hint == FCB_ENGINE_TENSOR && engine != ENG_UMMA) { set_error("engine_hint TENSOR: no tensor-core plan covers this layer"); return FCB_ERR_UNSUPPORTED; }
  if (engine != ENG_UMMA) {
    DirectParams& p = L->dp;
    const int deconv = g.kind == FCB_KIND_DECONV522;
    p.C = g.C; p.OFM = g.OFM; p.OFMp = (g.OFM + 63) / 64 * 64; p.KX = g.KX; p.KY = g.KY; p.IX = g.IX; p.IY = g.IY;
    p.OX = g.OX; p.OY = g.OY; p.SXe = deconv ? 1 : g.SX; p.SYe = deconv ? 1 : g.SY; p.PAD = g.pad_l; p.PADY = g.pad_u; p.deconv = deconv;
    p.in_bits = g.in_bits; p.in_signed = g.in_signed; p.in_word_bytes = (int)g.in_word_bytes; p.out_word_bytes = (int)g.out_word_bytes;
    p.out_x = g.out_x; p.out_y = g.out_y; p.tiles_x = (g.OX + 15) / 16; p.tiles_y = (g.OY + 7) / 8;
    p.DX = g.DX; p.DY = g.DY;
    p.patch_w = 15 * p.SXe + (g.KX - 1) * g.DX + 1; p.patch_h = 7 * p.SYe + (g.KY - 1) * g.DY + 1; p.mul_kind = g.weight_kind;
    p.in_img_bytes = g.in_img_bytes; p.out_img_bytes = g.out_img_bytes; p.epi = L->epi;
    const int cu = engine == ENG_XNOR ? g.C / 32 : g.C;
    const size_t budget = 96 * 1024;
    const bool imad2 = engine == ENG_IMAD && g.weight_kind != FCB_W_BINARY_XNOR;  // imad_conv_kernel: weights share the budget
    int cc = imad2 ? (int)(budget / ((size_t)p.patch_w * p.patch_h * 4 + (size_t)g.KX * g.KY * 128)) & ~3
                   : (int)(budget / ((size_t)p.patch_w * p.patch_h * 4));
    // weights that fit 8 bits (every reference config) run on the packed dot-product instructions: 4 MACs per instruction for lanes
    // of <= 8 bits, 2 for lanes of <= 16 bits (fcb_direct.cu, dot_conv_kernel)
    p.dot_pack = 0;
    if (imad2 && exp_int("FCB_DOT", 1)) {
      bool w8 = true;
      for (int32_t w : W) if (w < -128 || w > 127) { w8 = false; break; }
      const int pk = g.in_bits <= 8 ? 4 : 2;
      const int dcc = w8 ? dot_chunk_channels(p.patch_w, p.patch_h, g.KX * g.KY, g.C, pk, budget) : 0;
      if (dcc) { p.dot_pack = pk; cc = dcc; }
    }
    if (cc < 1) { set_error("kernel %dx%d stride %d: patch does not fit shared memory", g.KX, g.KY, g.SX); return FCB_ERR_UNSUPPORTED; }
    p.CC = std::min(cc, p.dot_pack ? cc : cu);
    L->smem = p.dot_pack ? dot_smem_bytes(p.patch_w, p.patch_h, g.KX * g.KY, p.CC, p.dot_pack)
              : imad2    ? imad_smem_bytes(p.patch_w, p.patch_h, g.KX * g.KY, p.CC)
                         : direct_smem_bytes(engine, p.patch_w, p.patch_h, p.CC);
    if (engine == ENG_XNOR) {
      const int KW = g.KX * g.KY * (g.C / 32);
      std::vector<uint32_t> Wb((size_t)KW * p.OFMp, 0u);
      for (int ch = 0; ch < g.OFM; ch++)
        for (int k = 0; k < g.K; k++) {
          const int tap = k / g.C, c = k % g.C;
          if (W[(size_t)ch * g.K + k]) Wb[(size_t)(tap * (g.C / 32) + c / 32) * p.OFMp + ch] |= 1u << (c & 31);
        }
      FCB_CUDA_OK(cudaMalloc(&L->d_wt, Wb.size() * 4));
      FCB_CUDA_OK(cudaMemcpy(L->d_wt, Wb.data(), Wb.size() * 4, cudaMemcpyHostToDevice));
    } else if (p.dot_pack) {
      const std::vector<uint32_t> Wd = dot_pack_weights(W, g.OFM, p.OFMp, g.C, g.KX * g.KY, p.dot_pack);
      FCB_CUDA_OK(cudaMalloc(&L->d_wt, Wd.size() * 4));
      FCB_CUDA_OK(cudaMemcpy(L->d_wt, Wd.data(), Wd.size() * 4, cudaMemcpyHostToDevice));
    } else {
      std::vector<int16_t> Wt((size_t)g.K * p.OFMp, 0);
      for (int ch = 0; ch < g.OFM; ch++)
        for (int k = 0; k < g.K; k++) Wt[(size_t)k * p.OFMp + ch] = (int16_t)W[(size_t)ch * g.K + k];
      FCB_CUDA_OK(cudaMalloc(&L->d_wt, Wt.size() * 2));
      FCB_CUDA_OK(cudaMemcpy(L->d_wt, Wt.data(), Wt.size() * 2, cudaMemcpyHostToDevice));
    }
    p.wt = L->d_wt;
  }
  if (L->lowered && (rc = lowered_reserve(L))) return rc;  // the lowering scratch is allocated here, never inside a run call
  if (L->post_pool && (rc = unpooled_reserve(L))) return rc;
  guard.armed = false;
  *out = L;
  return FCB_OK;
}

int fcb_layer_set_params(fcb_layer* L, const void* weights, const void* thresholds, const void* bias) {
  if (!L) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  // parameters are re-laid-out per plan (K-major tiles, sorted threshold tables, folded bias rows ...): build a fresh layer from the
  // stored descriptor and adopt its device state; the caller's handle, launch counter and staging slots stay
  fcb_layer* fresh = nullptr;
  int rc = fcb_layer_create(&L->desc, weights, thresholds, bias, L->device, &fresh);
  if (rc) return rc;
  std::swap(L->engine, fresh->engine);
  std::swap(L->d_wt, fresh->d_wt); std::swap(L->d_wt4, fresh->d_wt4); std::swap(L->d_bias, fresh->d_bias);
  std::swap(L->d_thr, fresh->d_thr); std::swap(L->d_thr_cm, fresh->d_thr_cm);
  std::swap(L->d_thr_lut, fresh->d_thr_lut); std::swap(L->d_thr_lo, fresh->d_thr_lo);
  std::swap(L->epi, fresh->epi); std::swap(L->dp, fresh->dp); std::swap(L->cw, fresh->cw); std::swap(L->smem, fresh->smem);
  std::swap(L->umma, fresh->umma);
  std::swap(L->d_unpooled[0], fresh->d_unpooled[0]); std::swap(L->d_unpooled[1], fresh->d_unpooled[1]); std::swap(L->unpooled_imgs, fresh->unpooled_imgs);
  std::swap(L->pool_cw, fresh->pool_cw); std::swap(L->post_pool, fresh->post_pool); std::swap(L->gi, fresh->gi);
  std::swap(L->lowered, fresh->lowered); std::swap(L->lower_bits, fresh->lower_bits); std::swap(L->ip, fresh->ip);
  std::swap(L->d_scratch[0], fresh->d_scratch[0]); std::swap(L->d_scratch[1], fresh->d_scratch[1]); std::swap(L->scratch_imgs, fresh->scratch_imgs); std::swap(L->scratch_img_bytes, fresh->scratch_img_bytes);
  memcpy(L->plan_desc, fresh->plan_desc, sizeof(L->plan_desc));
  fcb_layer_destroy(fresh);
  return FCB_OK;
}

int fcb_layer_set_param_stream(fcb_layer* L, const void* param_words, const void* thresholds, const void* bias) {
  if (!L || !param_words) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  // one period of the GenParamStream sequence (dma.h:214-236): word `tile` carries m_weights[pe][tile] in bits
  // [pe*SIMD*WP, (pe+1)*SIMD*WP) -- "Little Endian PE order" -- which Matrix_Vector_Activate_Stream_Batch slices back apart
  // (mvau.hpp:262-266).  Re-assemble the m_weights[PE][TILES] image and take the ordinary path.
  const auto& g = L->g;
  const uint32_t tiles = g.SF * g.NF, lane_bits = g.simd * g.w_bits;
  const size_t sw = word_bytes(lane_bits * g.pe);
  std::vector<uint8_t> img(g.weight_bytes, 0);
  const uint8_t* src = (const uint8_t*)param_words;
  for (uint32_t t = 0; t < tiles; t++)
    for (uint32_t pe = 0; pe < g.pe; pe++) {
      uint8_t* dst = img.data() + ((size_t)pe * tiles + t) * g.w_word_bytes;
      for (uint32_t b = 0; b < lane_bits; b += 8) {
        const uint32_t n = std::min<uint32_t>(8, lane_bits - b);
        dst[b >> 3] = (uint8_t)get_bits(src + (size_t)t * sw, (uint64_t)pe * lane_bits + b, n);
      }
    }
  return fcb_layer_set_params(L, img.data(), thresholds, bias);
}

const char* fcb_layer_engine(const fcb_layer* L) {
  if (!L) return "";
  return L->engine == ENG_UMMA ? "umma_i8" : L->engine == ENG_XNOR ? "xnor_popc" : L->engine == ENG_CHANWISE ? "chanwise" : "imad";
}
const char* fcb_layer_plan(const fcb_layer* L) {
  if (!L) return "";
  if (L->lowered) return L->plan_desc;
  if (L->engine == ENG_CHANWISE) {
    const int v = chanwise_vector_words(L->cw);
    return v == 4   ? "channel-wise streaming unit, byte lanes: thread = 16 channels of an output pixel (16-byte loads per tap)"
           : v == 1 ? "channel-wise streaming unit, byte lanes: thread = 4 channels of an output pixel (4-byte loads per tap)"
                    : "channel-wise streaming unit: warp = output pixel, lanes walk the channels";
  }
  if (L->engine == ENG_UMMA) return umma_plan_describe(L->umma);
  if (L->dp.dot_pack == 4) return "direct 16x8-pixel x 64-channel CTA tiles, IDP.4A (4 channels per patch word, weights as bytes in shared memory)";
  if (L->dp.dot_pack == 2) return "direct 16x8-pixel x 64-channel CTA tiles, IDP.2A (2 channels per patch word, weights as bytes in shared memory)";
  return "direct 16x8-pixel x 64-channel CTA tiles (patch and weight pairs in shared memory)";
}
uint64_t fcb_layer_launches(const fcb_layer* L) { return L ? L->launches : 0; }

}  // extern "C"

// Scratch of a lowered layer (im2col rows / expanded bits): two slots of <= 64 MiB (at least 16 images' worth of alignment), one per
// staging slot of the host-buffer call, so the lowering kernel of chunk k+1 never overwrites rows the GEMM of chunk k still reads.
static int lowered_reserve(fcb_layer* L) {
  size_t cap = std::max<size_t>(1, ((size_t)64 << 20) / L->scratch_img_bytes);
  cap = align_chunk(cap, L->gi.in_img_bytes, L->gi.out_img_bytes);
  for (int i = 0; i < 2; i++) FCB_CUDA_OK(cudaMalloc(&L->d_scratch[i], L->scratch_img_bytes * cap));
  L->scratch_imgs = cap;
  return FCB_OK;
}

static int unpooled_reserve(fcb_layer* L) {
  size_t cap = std::max<size_t>(1, ((size_t)128 << 20) / L->gi.out_img_bytes);
  cap = align_chunk(cap, L->g.in_img_bytes, L->g.out_img_bytes);
  for (int i = 0; i < 2; i++) FCB_CUDA_OK(cudaMalloc(&L->d_unpooled[i], L->gi.out_img_bytes * cap));
  L->unpooled_imgs = cap;
  return FCB_OK;
}

static int engine_run_device(fcb_layer* L, const void* d_in, void* d_out, uint32_t numReps, cudaStream_t st, int slot);

static int layer_run_device(fcb_layer* L, const void* d_in, void* d_out, uint32_t numReps, cudaStream_t st, int slot) {
  if (!L->post_pool) return engine_run_device(L, d_in, d_out, numReps, st, slot);
  for (size_t n0 = 0; n0 < numReps; n0 += L->unpooled_imgs) {
    const uint32_t nb = (uint32_t)std::min<size_t>(L->unpooled_imgs, numReps - n0);
    int rc = engine_run_device(L, (const uint8_t*)d_in + n0 * L->g.in_img_bytes, L->d_unpooled[slot], nb, st, slot);
    if (rc) return rc;
    ChanParams c = L->pool_cw;
    c.in = (const uint8_t*)L->d_unpooled[slot];
    c.out = (uint8_t*)d_out + n0 * L->g.out_img_bytes;
    if (L->g.out_word_bytes * 8 != (size_t)L->g.OFM * L->g.out_bits) FCB_CUDA_OK(cudaMemsetAsync(c.out, 0, L->g.out_img_bytes * nb, st));
    rc = launch_chanwise(c, (int)nb, st);
    if (rc) return rc;
    L->launches += (nb + 65534) / 65535;
  }
  return FCB_OK;
}

// the layer's engine on `numReps` images (geometry L->gi: the layer without a separately-run pool)
static int engine_run_device(fcb_layer* L, const void* d_in, void* d_out, uint32_t numReps, cudaStream_t st, int slot) {
  if (L->lowered) {
    for (size_t n0 = 0; n0 < numReps; n0 += L->scratch_imgs) {
      const int nb = (int)std::min<size_t>(L->scratch_imgs, numReps - n0);
      Im2colParams ip = L->ip;
      ip.in = (const uint8_t*)d_in + n0 * L->gi.in_img_bytes;
      ip.out = (uint8_t*)L->d_scratch[slot];
      int rc = L->lower_bits ? launch_expand_bits(ip, nb, st) : launch_im2col(ip, nb, st);
      if (rc) return rc;
      L->launches++;
      rc = umma_run(L->umma, L->d_scratch[slot], (uint8_t*)d_out + n0 * L->gi.out_img_bytes, nb, st, &L->launches);
      if (rc) return rc;
    }
    return FCB_OK;
  }
  if (L->engine == ENG_UMMA) return umma_run(L->umma, d_in, d_out, (int)numReps, st, &L->launches);
  // containers with padding bits (e.g. ap_uint<24> in 4 bytes): writers zero them
  if (L->gi.out_word_bytes * 8 != (size_t)L->gi.OFM * L->gi.out_bits) {
    FCB_CUDA_OK(cudaMemsetAsync(d_out, 0, L->gi.out_img_bytes * numReps, st));
  }
  if (L->engine == ENG_CHANWISE) {
    ChanParams c = L->cw;
    c.in = (const uint8_t*)d_in;
    c.out = (uint8_t*)d_out;
    int rc = launch_chanwise(c, (int)numReps, st);
    if (rc == FCB_OK) L->launches += (numReps + 65534) / 65535;
    return rc;
  }
  DirectParams p = L->dp;
  p.in = (const uint8_t*)d_in;
  p.out = (uint8_t*)d_out;
  int rc = launch_direct(p, L->engine, (int)numReps, L->smem, st);
  if (rc == FCB_OK) L->launches += (numReps + 65534) / 65535;
  return rc;
}

extern "C" {

int fcb_layer_run_device(fcb_layer* L, const void* d_in, void* d_out, uint32_t numReps, void* stream) {
  if (!L || !d_in || !d_out) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  if (numReps == 0) return FCB_OK;
  FCB_ON_DEVICE(L->device);
  return layer_run_device(L, d_in, d_out, numReps, (cudaStream_t)stream, 0);
}

int fcb_layer_set_host_chunk(fcb_layer* L, uint32_t images) {
  if (!L) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  L->host_chunk = images;
  return FCB_OK;
}

int fcb_layer_run(fcb_layer* L, const void* in_words, void* out_words, uint32_t numReps) {
  if (!L || !in_words || !out_words) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  if (numReps == 0) return FCB_OK;
  FCB_ON_DEVICE(L->device);
  // chunk so that one slot stays <= 256 MiB of input or output; two slots overlap copy and compute
  size_t chunk = L->host_chunk ? L->host_chunk : std::max<size_t>(1, ((size_t)256 << 20) / std::max(L->g.in_img_bytes, L->g.out_img_bytes));
  chunk = align_chunk(std::min<size_t>(chunk, numReps), L->g.in_img_bytes, L->g.out_img_bytes);
  if (L->s_imgs < chunk) {
    L->s_imgs = 0;  // a failed reallocation must not leave a stale capacity behind
    for (int i = 0; i < 2; i++) {
      cudaFree(L->s_in[i]); cudaFree(L->s_out[i]);
      L->s_in[i] = L->s_out[i] = nullptr;
    }
    for (int i = 0; i < 2; i++) {
      FCB_CUDA_OK(cudaMalloc(&L->s_in[i], L->g.in_img_bytes * chunk));
      FCB_CUDA_OK(cudaMalloc(&L->s_out[i], L->g.out_img_bytes * chunk));
      if (!L->s_stream[i]) FCB_CUDA_OK(cudaStreamCreateWithFlags(&L->s_stream[i], cudaStreamNonBlocking));
    }
    L->s_imgs = chunk;
  }
  int slot = 0;
  for (size_t n0 = 0; n0 < numReps; n0 += chunk, slot ^= 1) {
    const size_t nb = std::min<size_t>(chunk, numReps - n0);
    cudaStream_t st = L->s_stream[slot];
    FCB_CUDA_OK(cudaMemcpyAsync(L->s_in[slot], (const uint8_t*)in_words + n0 * L->g.in_img_bytes, nb * L->g.in_img_bytes, cudaMemcpyHostToDevice, st));
    int rc = layer_run_device(L, L->s_in[slot], L->s_out[slot], (uint32_t)nb, st, slot);
    if (rc) return rc;
    FCB_CUDA_OK(cudaMemcpyAsync((uint8_t*)out_words + n0 * L->g.out_img_bytes, L->s_out[slot], nb * L->g.out_img_bytes, cudaMemcpyDeviceToHost, st));
  }
  FCB_CUDA_OK(cudaStreamSynchronize(L->s_stream[0]));
  FCB_CUDA_OK(cudaStreamSynchronize(L->s_stream[1]));
  return FCB_OK;
}

// ---- layer chain ------------------------------------------------------------------------
int fcb_net_create(fcb_layer* const* layers, uint32_t n, fcb_net** out) {
  if (!layers || !n || !out) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  for (uint32_t i = 0; i < n; i++) {
    if (!layers[i]) { set_error("layer %u is NULL", i); return FCB_ERR_INVALID_ARG; }
    if (layers[i]->device != layers[0]->device) { set_error("layers live on different devices"); return FCB_ERR_INVALID_ARG; }
    if (i + 1 < n) {
      const Geom &a = layers[i]->g, &b = layers[i + 1]->g;
      if (a.OFM * a.out_bits != b.C * b.in_bits || a.out_x != b.IX || a.out_y != b.IY) {
        set_error("layer %u output (%dx%d, %d bits) does not feed layer %u input (%dx%d, %d bits)", i, a.out_x, a.out_y,
                  a.OFM * a.out_bits, i + 1, b.IX, b.IY, b.C * b.in_bits);
        return FCB_ERR_SHAPE;
      }
    }
  }
  fcb_net* N = new fcb_net();
  N->layers.assign(layers, layers + n);
  N->bufs.assign(n > 1 ? n - 1 : 0, nullptr);
  N->device = layers[0]->device;
  *out = N;
  return FCB_OK;
}

void fcb_net_destroy(fcb_net* N) {
  if (!N) return;
  DeviceScope ds(N->device);
  for (void* b : N->bufs) cudaFree(b);
  for (int i = 0; i < 2; i++) {
    cudaFree(N->s_in[i]); cudaFree(N->s_out[i]);
    if (N->ev_in[i]) cudaEventDestroy(N->ev_in[i]);
    if (N->ev_run[i]) cudaEventDestroy(N->ev_run[i]);
    if (N->ev_out[i]) cudaEventDestroy(N->ev_out[i]);
  }
  if (N->st_in) cudaStreamDestroy(N->st_in);
  if (N->st_run) cudaStreamDestroy(N->st_run);
  if (N->st_out) cudaStreamDestroy(N->st_out);
  delete N;
}

static int net_reserve(fcb_net* N, size_t imgs) {
  if (N->cap_imgs >= imgs) return FCB_OK;
  N->cap_imgs = 0;  // a failed reallocation must not leave a stale capacity behind
  for (size_t i = 0; i < N->bufs.size(); i++) {
    cudaFree(N->bufs[i]);
    N->bufs[i] = nullptr;
  }
  for (size_t i = 0; i < N->bufs.size(); i++) FCB_CUDA_OK(cudaMalloc(&N->bufs[i], N->layers[i]->g.out_img_bytes * imgs));
  N->cap_imgs = imgs;
  return FCB_OK;
}

}  // extern "C"

static int net_run_device(fcb_net* N, const void* d_in, void* d_out, uint32_t numReps, cudaStream_t st, int slot) {
  // bound the intermediates: process in chunks of images (default: the largest intermediate stays <= 1 GiB)
  size_t biggest = 1;
  for (size_t i = 0; i + 1 < N->layers.size(); i++) biggest = std::max(biggest, N->layers[i]->g.out_img_bytes);
  size_t chunk = N->device_chunk ? N->device_chunk : std::max<size_t>(1, ((size_t)1 << 30) / biggest);
  const size_t in_b = N->layers.front()->g.in_img_bytes, out_b = N->layers.back()->g.out_img_bytes;
  chunk = align_chunk(std::min<size_t>(chunk, numReps), in_b, out_b);
  int rc = net_reserve(N, chunk);
  if (rc) return rc;
  for (size_t n0 = 0; n0 < numReps; n0 += chunk) {
    const uint32_t nb = (uint32_t)std::min<size_t>(chunk, numReps - n0);
    const void* src = (const uint8_t*)d_in + n0 * in_b;
    for (size_t i = 0; i < N->layers.size(); i++) {
      void* dst = (i + 1 == N->layers.size()) ? (void*)((uint8_t*)d_out + n0 * out_b) : N->bufs[i];
      rc = layer_run_device(N->layers[i], src, dst, nb, st, slot);
      if (rc) return rc;
      src = dst;
    }
  }
  return FCB_OK;
}

extern "C" {

int fcb_net_run_device(fcb_net* N, const void* d_in, void* d_out, uint32_t numReps, void* stream) {
  if (!N || !d_in || !d_out) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  if (!numReps) return FCB_OK;
  FCB_ON_DEVICE(N->device);
  return net_run_device(N, d_in, d_out, numReps, (cudaStream_t)stream, 0);
}

int fcb_net_set_host_chunk(fcb_net* N, uint32_t images) {
  if (!N) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  N->host_chunk = images;
  return FCB_OK;
}

int fcb_net_set_device_chunk(fcb_net* N, uint32_t images) {
  if (!N) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  N->device_chunk = images;
  return FCB_OK;
}

int fcb_net_run(fcb_net* N, const void* in_words, void* out_words, uint32_t numReps) {
  if (!N || !in_words || !out_words) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  if (!numReps) return FCB_OK;
  FCB_ON_DEVICE(N->device);
  const size_t in_b = N->layers.front()->g.in_img_bytes, out_b = N->layers.back()->g.out_img_bytes;
  // chunks of <= 64 MiB of input or output: small enough that the pipeline fills quickly, large enough to hide launch overheads
  size_t chunk = N->host_chunk ? N->host_chunk : std::max<size_t>(1, ((size_t)64 << 20) / std::max(in_b, out_b));
  chunk = align_chunk(std::min<size_t>(chunk, numReps), in_b, out_b);
  if (N->s_imgs < chunk) {
    N->s_imgs = 0;  // a failed reallocation must not leave a stale capacity behind
    for (int i = 0; i < 2; i++) {
      cudaFree(N->s_in[i]); cudaFree(N->s_out[i]);
      N->s_in[i] = N->s_out[i] = nullptr;
    }
    for (int i = 0; i < 2; i++) {
      FCB_CUDA_OK(cudaMalloc(&N->s_in[i], in_b * chunk));
      FCB_CUDA_OK(cudaMalloc(&N->s_out[i], out_b * chunk));
      if (!N->ev_in[i]) FCB_CUDA_OK(cudaEventCreateWithFlags(&N->ev_in[i], cudaEventDisableTiming));
      if (!N->ev_run[i]) FCB_CUDA_OK(cudaEventCreateWithFlags(&N->ev_run[i], cudaEventDisableTiming));
      if (!N->ev_out[i]) FCB_CUDA_OK(cudaEventCreateWithFlags(&N->ev_out[i], cudaEventDisableTiming));
    }
    if (!N->st_in) FCB_CUDA_OK(cudaStreamCreateWithFlags(&N->st_in, cudaStreamNonBlocking));
    if (!N->st_run) FCB_CUDA_OK(cudaStreamCreateWithFlags(&N->st_run, cudaStreamNonBlocking));
    if (!N->st_out) FCB_CUDA_OK(cudaStreamCreateWithFlags(&N->st_out, cudaStreamNonBlocking));
    N->s_imgs = chunk;
  }
  size_t k = 0;
  for (size_t n0 = 0; n0 < numReps; n0 += chunk, k++) {
    const size_t nb = std::min(chunk, (size_t)numReps - n0);
    const int slot = (int)(k & 1);
    // H2D of chunk k: its input slot was last read by the layers of chunk k-2
    if (k >= 2) FCB_CUDA_OK(cudaStreamWaitEvent(N->st_in, N->ev_run[slot], 0));
    FCB_CUDA_OK(cudaMemcpyAsync(N->s_in[slot], (const uint8_t*)in_words + n0 * in_b, nb * in_b, cudaMemcpyHostToDevice, N->st_in));
    FCB_CUDA_OK(cudaEventRecord(N->ev_in[slot], N->st_in));
    // layers of chunk k: need its input, and its output slot drained by the D2H of chunk k-2.  All chunks run on ONE stream, so the
    // intermediates and lowering scratch (slot 0) are never shared between chunks in flight.
    FCB_CUDA_OK(cudaStreamWaitEvent(N->st_run, N->ev_in[slot], 0));
    if (k >= 2) FCB_CUDA_OK(cudaStreamWaitEvent(N->st_run, N->ev_out[slot], 0));
    int rc = net_run_device(N, N->s_in[slot], N->s_out[slot], (uint32_t)nb, N->st_run, 0);
    if (rc) return rc;
    FCB_CUDA_OK(cudaEventRecord(N->ev_run[slot], N->st_run));
    // D2H of chunk k
    FCB_CUDA_OK(cudaStreamWaitEvent(N->st_out, N->ev_run[slot], 0));
    FCB_CUDA_OK(cudaMemcpyAsync((uint8_t*)out_words + n0 * out_b, N->s_out[slot], nb * out_b, cudaMemcpyDeviceToHost, N->st_out));
    FCB_CUDA_OK(cudaEventRecord(N->ev_out[slot], N->st_out));
  }
  FCB_CUDA_OK(cudaStreamSynchronize(N->st_out));
  FCB_CUDA_OK(cudaStreamSynchronize(N->st_run));
  return FCB_OK;
}

uint64_t fcb_net_launches(const fcb_net* N) {
  uint64_t s = 0;
  if (N) for (auto* l : N->layers) s += l->launches;
  return s;
}

static int add_check(const fcb_add_desc* d) {
  if (!d || d->struct_size != sizeof(fcb_add_desc)) { set_error("bad fcb_add_desc"); return FCB_ERR_INVALID_ARG; }
  if (!d->channels || d->in1_bits < 1 || d->in1_bits > 32 || d->in2_bits < 1 || d->in2_bits > 32 || d->out_bits < 1 || d->out_bits > 32) {
    set_error("AddStreams: lane widths 1..32, channels > 0"); return FCB_ERR_UNSUPPORTED;
  }
  return FCB_OK;
}

int fcb_add_streams_device(const fcb_add_desc* d, const void* d_in1, const void* d_in2, void* d_out, uint64_t n_words, int device, void* stream) {
  int rc = add_check(d);
  if (rc) return rc;
  if (!d_in1 || !d_in2 || !d_out) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  FCB_ON_DEVICE(device);
  return launch_add_streams(d_in1, d_in2, d_out, n_words, (int)d->channels, (int)d->in1_bits, d->in1_signed ? 1 : 0, (int)d->in2_bits,
                            d->in2_signed ? 1 : 0, (int)d->out_bits, d->offset, (int)word_bytes(d->channels * d->in1_bits),
                            (int)word_bytes(d->channels * d->in2_bits), (int)word_bytes(d->channels * d->out_bits), (cudaStream_t)stream);
}

int fcb_add_streams(const fcb_add_desc* d, const void* in1, const void* in2, void* out, uint64_t n_words, int device) {
  int rc = add_check(d);
  if (rc) return rc;
  if (!in1 || !in2 || !out) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  if (!n_words) return FCB_OK;
  FCB_ON_DEVICE(device);
  const size_t b1 = word_bytes(d->channels * d->in1_bits) * n_words, b2 = word_bytes(d->channels * d->in2_bits) * n_words,
               bo = word_bytes(d->channels * d->out_bits) * n_words;
  struct Bufs {
    void *a = nullptr, *b = nullptr, *o = nullptr;
    ~Bufs() { cudaFree(a); cudaFree(b); cudaFree(o); }
  } B;
  FCB_CUDA_OK(cudaMalloc(&B.a, b1));
  FCB_CUDA_OK(cudaMalloc(&B.b, b2));
  FCB_CUDA_OK(cudaMalloc(&B.o, bo));
  FCB_CUDA_OK(cudaMemcpy(B.a, in1, b1, cudaMemcpyHostToDevice));
  FCB_CUDA_OK(cudaMemcpy(B.b, in2, b2, cudaMemcpyHostToDevice));
  rc = fcb_add_streams_device(d, B.a, B.b, B.o, n_words, device, nullptr);
  if (rc) return rc;
  FCB_CUDA_OK(cudaMemcpy(out, B.o, bo, cudaMemcpyDeviceToHost));
  return FCB_OK;
}

int fcb_synth_fill(void* d_ptr, size_t n_bytes, uint64_t seed, uint32_t mask, uint64_t offset, void* stream) {
  if (!d_ptr) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  return synth_fill(d_ptr, n_bytes, seed, mask, offset, (cudaStream_t)stream);
}

}  // extern "C"
