// fcb_plan.cu -- host side of the "umma_i8" engine: eligibility, weight re-layout into the K-major s8 operand images the
// tcgen05 kernels of fcb_umma2.cu read, tensor-map encoding, plan ownership.  No kernels here.
//
//   D[ch][pixel] = sum_k W[ch][k] * A[pixel][k],  k = (ky*Kx + kx)*C + c   (mvau.hpp:122-178, window order
//   slidingwindow.h:1302-1313), followed by the fused activation stage (fcb_epilogue.cuh).
#include <string.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "fcb_umma_common.h"

namespace fcb {

struct UmmaPlan {
  Geom g;
  int device = 0;
  int8_t* d_w = nullptr;          // operand image of the weights (layout depends on the plan family)
  int8_t* d_zero_bias = nullptr;  // thin-input plans with the bias folded into the weights
  int num_sms = 148;
  Umma2Plan* v2 = nullptr;        // the plan proper (fcb_umma2.cu)
  UmmaV1* v1 = nullptr;           // experiment builds only (fcb_umma_v1.cu)
  char desc[224] = "";
};

void umma_plan_destroy(UmmaPlan* P) {
  if (!P) return;
  if (P->v2) umma2_plan_destroy(P->v2);
#ifdef FCB_EXPERIMENT
  if (P->v1) umma_v1_destroy(P->v1);
#endif
  cudaFree(P->d_w);
  cudaFree(P->d_zero_bias);
  delete P;
}

namespace {
// a half-built plan is released on every early return (FCB_CUDA_OK included); release() hands it to the caller
struct PlanHolder {
  UmmaPlan* P;
  explicit PlanHolder(const Geom& g, int device) : P(new UmmaPlan()) { P->g = g; P->device = device; }
  ~PlanHolder() { umma_plan_destroy(P); }
  UmmaPlan* operator->() { return P; }
  UmmaPlan* release() { UmmaPlan* r = P; P = nullptr; return r; }
};
int upload(int8_t** dst, const std::vector<int8_t>& src) {
  FCB_CUDA_OK(cudaMalloc(dst, src.size()));
  FCB_CUDA_OK(cudaMemcpy(*dst, src.data(), src.size(), cudaMemcpyHostToDevice));
  return FCB_OK;
}
int query_sms(int device, int* n) {
  FCB_CUDA_OK(cudaDeviceGetAttribute(n, cudaDevAttrMultiProcessorCount, device));
  return FCB_OK;
}
}  // namespace

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
int umma_encode_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  return umma_encode_map_ex(m, base, 1, 128, rank, dims, strides_bytes, box);
}
// elem_bytes: 1 (u8) or 4 (u32 elements: boxes wider than 256 bytes); swizzle: 0 (dense box image) or 128
int umma_encode_map_ex(CUtensorMap* m, void* base, int elem_bytes, int swizzle, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return FCB_ERR_CUDA; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) gs[i] = strides_bytes[i];
  CUresult r = enc(m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, base, gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank); return FCB_ERR_CUDA; }
  return FCB_OK;
}

int umma_eligible(const Geom& g) {
  if (g.weight_kind != FCB_W_FIXED || g.w_bits > 8 || g.in_bits != 8) return 0;
  if (g.in_word_bytes != (size_t)g.C) return 0;              // stream image == dense NHWC bytes
  if (g.C % 16) return 0;                                    // TMA strides are multiples of 16 bytes
  if (g.OFM > 256) return 0;
  if (g.KX * g.KY > 32) return 0;
  if ((uint64_t)g.K * 255ull * 128ull >= (1ull << 31)) return 0;  // exact int32 accumulation
  if (g.pool > 2) return 0;
  if (g.kind == FCB_KIND_DECONV522) return g.pool == 1;  // (its padding is fixed: derive_geom)
  if (g.SX != g.SY) return 0;
  // stride 1: a partial last channel chunk is zero-filled by TMA (it then multiplies the next tap's weights by 0);
  // stride 2: the parity view packs two pixels per row, so chunks must not straddle pixels
  if (g.SX == 1) return 1;
  if (g.SX == 2) return (g.C % 128 == 0) && (g.IX % 2 == 0) && (g.IY % 2 == 0);
  return 0;
}

int umma_plan_create(const Geom& g, const std::vector<int32_t>& W, const EpiParams& epi, int device, UmmaPlan** out) {
  *out = nullptr;
  PlanHolder P(g, device);
  int rc = query_sms(device, &P->num_sms);
  if (rc) return rc;
  // weights: s8 [OFM padded to whole 128-channel blocks][K], K contiguous (k = (ky*KX+kx)*C + c) -- the A operand.  Rows are
  // padded with zeros in memory: an all-out-of-bounds TMA box row costs as much as a real one, a zero row in memory does not
  const int rows_pad = (g.OFM + 127) / 128 * 128;
  std::vector<int8_t> w8((size_t)rows_pad * g.K, 0);
  for (size_t i = 0; i < (size_t)g.OFM * g.K; i++) w8[i] = (int8_t)W[i];
  if ((rc = upload(&P->d_w, w8))) return rc;
#ifdef FCB_EXPERIMENT
  const char* v1only = exp_env("FCB_UMMA_V1");
  if (v1only && v1only[0] == '1' && g.pad_l == g.pad_r && g.pad_u == g.pad_d && g.pad_l == g.pad_u && g.DX == 1 && g.DY == 1) {  // (v1: symmetric padding only)
    rc = umma_v1_create(g, P->d_w, epi, P->num_sms, &P->v1);
    if (rc) return rc;
    snprintf(P->desc, sizeof(P->desc), "v1 per-tap TMA");
    *out = P.release();
    return FCB_OK;
  }
#endif
  rc = umma2_plan_create(g, P->d_w, epi, P->num_sms, &P->v2);
  if (rc == FCB_ERR_UNSUPPORTED) set_error("no tensor-core plan for this shape");
  if (rc) return rc;
  umma2_describe(P->v2, P->desc, sizeof(P->desc));
  *out = P.release();
  return FCB_OK;
}

// Thin-input layers (one 4-byte word per pixel): W4 is [OFM][128], k = (ky*KX + kx)*4 + lane.
int umma_plan_create_thin(const Geom& g, const std::vector<int32_t>& W4, const EpiParams& epi, const int8_t* bias_host, int device, UmmaPlan** out) {
  *out = nullptr;
  PlanHolder P(g, device);
  int rc = query_sms(device, &P->num_sms);
  if (rc) return rc;
  const int rows_pad = (g.OFM + 127) / 128 * 128;
  std::vector<int8_t> w8((size_t)rows_pad * 128, 0);
  for (size_t i = 0; i < (size_t)g.OFM * 128; i++) w8[i] = (int8_t)W4[i];
  // bias + ReLU on the wrapped 8-bit lane (conv_nonsquare_top.cpp:267-278): ((acc mod 256) + bias) mod 256 = (acc + bias) mod 256, so
  // the bias can ride in the GEMM as the weight of a constant-1 activation in the first unused window word
  const int nw = g.KX * g.KY;
  int bias_word = -1;
  if (bias_host && epi.act_kind == FCB_ACT_BIAS_RELU && epi.out_bits == 8 && epi.acc_bits == 8 && (nw == 25 || nw == 9) &&
      (uint64_t)(g.K + 1) * 255ull * 128ull < (1ull << 31) && !exp_env("FCB_U2_NO_FOLD")) {
    bias_word = nw;
    for (int ch = 0; ch < g.OFM; ch++) w8[(size_t)ch * 128 + 4 * nw] = bias_host[ch];
  }
  if ((rc = upload(&P->d_w, w8))) return rc;
  EpiParams epi2 = epi;
  if (bias_word >= 0) {  // every epilogue variant now adds a zero bias
    FCB_CUDA_OK(cudaMalloc(&P->d_zero_bias, rows_pad));
    FCB_CUDA_OK(cudaMemset(P->d_zero_bias, 0, rows_pad));
    epi2.bias = P->d_zero_bias;
  }
  rc = umma2_plan_create_thin(g, P->d_w, epi2, P->num_sms, bias_word, &P->v2);
  if (rc) return rc;
  umma2_describe(P->v2, P->desc, sizeof(P->desc));
  *out = P.release();
  return FCB_OK;
}

// Thin-output transposed conv (deconv522, OFM 3..4): weights regrouped by input shift.  Output phase (py, px) uses tap
// (ky, kx) = (2*offy + 2 - py, 2*offx + 2 - px) at shift (offy, offx) in {-1,0,1}^2 when that tap exists.
int umma_plan_create_dthin(const Geom& g, const std::vector<int32_t>& W, const EpiParams& epi, int device, UmmaPlan** out) {
  *out = nullptr;
  if (g.kind != FCB_KIND_DECONV522 || g.OFM < 3 || g.OFM > 4 || g.C % 128 || g.C > 256 || exp_env("FCB_U2_NO_DTHIN")) return FCB_ERR_UNSUPPORTED;
  PlanHolder P(g, device);
  int rc = query_sms(device, &P->num_sms);
  if (rc) return rc;
  const int cch = g.C / 128;
  if (!exp_env("FCB_U2_NO_DCOL")) {
    // col2im form: one GEMM over rows (tap word * 4 + channel, dcol_word()), K = the input channels; the taps are summed after the GEMM
    std::vector<int8_t> wc((size_t)cch * 128 * 128, 0);
    for (int cc = 0; cc < cch; cc++)
      for (int t = 0; t < 25; t++)
        for (int o = 0; o < g.OFM; o++)
          for (int c = 0; c < 128; c++)
            wc[((size_t)cc * 128 + dcol_word(t) * 4 + o) * 128 + c] = (int8_t)W[(size_t)o * g.K + t * g.C + cc * 128 + c];
    if ((rc = upload(&P->d_w, wc))) return rc;
    rc = umma2_plan_create_dcol(g, P->d_w, epi, P->num_sms, &P->v2);
    if (rc == FCB_OK) {
      umma2_describe(P->v2, P->desc, sizeof(P->desc));
      *out = P.release();
      return FCB_OK;
    }
    cudaFree(P->d_w);
    P->d_w = nullptr;
    if (rc != FCB_ERR_UNSUPPORTED) return rc;
  }
#ifdef FCB_EXPERIMENT
  // 9 shift blocks x N = 16 (pixels on M, no col2im): kept as an independent cross-check of the col2im form
  std::vector<int8_t> w8((size_t)9 * cch * 16 * 128, 0);
  for (int cc = 0; cc < cch; cc++)
    for (int sft = 0; sft < 9; sft++) {
      const int offy = sft / 3 - 1, offx = sft % 3 - 1;
      for (int ph = 0; ph < 4; ph++) {
        const int ky = 2 * offy + 2 - ph / 2, kx = 2 * offx + 2 - ph % 2;
        if (ky < 0 || ky > 4 || kx < 0 || kx > 4) continue;
        for (int o = 0; o < g.OFM; o++)
          for (int c = 0; c < 128; c++)
            w8[(((size_t)(cc * 9 + sft) * 16) + ph * 4 + o) * 128 + c] = (int8_t)W[(size_t)o * g.K + (ky * 5 + kx) * g.C + cc * 128 + c];
      }
    }
  if ((rc = upload(&P->d_w, w8))) return rc;
  rc = umma2_plan_create_dthin(g, P->d_w, epi, P->num_sms, &P->v2);
  if (rc) return rc;
  umma2_describe(P->v2, P->desc, sizeof(P->desc));
  *out = P.release();
  return FCB_OK;
#else
  return FCB_ERR_UNSUPPORTED;
#endif
}

const char* umma_plan_describe(const UmmaPlan* P) { return P ? P->desc : ""; }

int umma_run(UmmaPlan* P, const void* d_in, void* d_out, int n_images, cudaStream_t st, uint64_t* launches) {
  if (((uintptr_t)d_in & 15) || ((uintptr_t)d_out & 15)) { set_error("device buffers must be 16-byte aligned"); return FCB_ERR_INVALID_ARG; }
  int rc;
#ifdef FCB_EXPERIMENT
  if (P->v1) rc = umma_v1_run(P->v1, d_in, d_out, n_images, st);
  else
#endif
    rc = umma2_run(P->v2, d_in, d_out, n_images, st);
  if (rc == FCB_OK && launches) (*launches)++;
  return rc;
}

}  // namespace fcb
