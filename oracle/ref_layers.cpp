// oracle/ref_layers.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Builds the REFERENCE's own templates (conv2d<> / deconv522<> from
// conv_nonsquare_top.cpp:71-280, and the library path
// FMPadding_nonsquare + StreamingDataWidthConverter_Batch +
// ConvolutionInputGenerator_NonSquare + Matrix_Vector_Activate_Batch +
// ThresholdsActivation / BinaryWeights / StreamingMaxPool*) into a shared
// library with a plain C interface, compiled from the sources where they lie
// under /root/reference against oracle/shim/.  Output: oracle/_ref/libref_layers.so
// (git-ignored).  Nothing of the reference is copied: this file only
// #includes it.
//
// Every entry point speaks the same packed byte images as include/finnconv_b200.h
// (stream words / m_weights[PE][TILES] / m_thresholds[PE][NF][NumTH] /
// bias m_weights[1][OFM] in "ap-word containers": 1,2,4,8 bytes for widths
// <=8,16,32,64, else 8*ceil(W/64), little-endian), so the reference, the C
// restatement (finn_oracle.c) and the CUDA library can be fed identical bytes.
//
// The 2.4 MB memdata_nonsquare.h is skipped here (guard PARAMS_HPP,
// memdata_nonsquare.h:1-2) and PARAM:: is declared empty: weights are
// caller-supplied.  The fixture weights are used by oracle/ref_net.cpp.
#include <cstdint>
#include <cstring>
#include <chrono>

#define AP_INT_MAX_W 16384
#include <hls_stream.h>
#include "ap_int.h"
#include "weights.hpp"
#include "config_nonsquare.h"

#define PARAMS_HPP
namespace PARAM {
static FixedPointWeights<CONV_0_SIMD, ap_int<CONV_0_W_BIT>, CONV_0_PE, CONV_0_W_TILES> weights_layer0;
static FixedPointWeights<CONV_1_SIMD, ap_int<CONV_1_W_BIT>, CONV_1_PE, CONV_1_W_TILES> weights_layer1;
static FixedPointWeights<CONV_2_SIMD, ap_int<CONV_2_W_BIT>, CONV_2_PE, CONV_2_W_TILES> weights_layer2;
static FixedPointWeights<CONV_3_SIMD, ap_int<CONV_3_W_BIT>, CONV_3_PE, CONV_3_W_TILES> weights_layer3;
static FixedPointWeights<CONV_4_SIMD, ap_int<CONV_4_W_BIT>, CONV_4_PE, CONV_4_W_TILES> weights_layer4;
static FixedPointWeights<CONV_5_SIMD, ap_int<CONV_5_W_BIT>, CONV_5_PE, CONV_5_W_TILES> weights_layer5;
static FixedPointWeights<CONV_6_SIMD, ap_int<CONV_6_W_BIT>, CONV_6_PE, CONV_6_W_TILES> weights_layer6;
static FixedPointWeights<CONV_7_SIMD, ap_int<CONV_7_W_BIT>, CONV_7_PE, CONV_7_W_TILES> weights_layer7;
static FixedPointWeights<1, ap_int<8>, 1, CONV_0_OFM_CH> bias_layer0;
static FixedPointWeights<1, ap_int<8>, 1, CONV_1_OFM_CH> bias_layer1;
static FixedPointWeights<1, ap_int<8>, 1, CONV_2_OFM_CH> bias_layer2;
static FixedPointWeights<1, ap_int<8>, 1, CONV_3_OFM_CH> bias_layer3;
static FixedPointWeights<1, ap_int<8>, 1, CONV_4_OFM_CH> bias_layer4;
static FixedPointWeights<1, ap_int<8>, 1, CONV_5_OFM_CH> bias_layer5;
static FixedPointWeights<1, ap_int<8>, 1, CONV_6_OFM_CH> bias_layer6;
static FixedPointWeights<1, ap_int<8>, 1, CONV_7_OFM_CH> bias_layer7;
}  // namespace PARAM

#include "conv_nonsquare_top.cpp"  // the reference top, unmodified (templates conv2d<>, deconv522<>)
#include "pool.hpp"                // pool functions of Pool_batch (not pulled in by bnn-library.h)

namespace {

constexpr int container_bytes(int w) { return w <= 8 ? 1 : w <= 16 ? 2 : w <= 32 ? 4 : w <= 64 ? 8 : 8 * ((w + 63) / 64); }

template <int W> ap_uint<W> load_word(const uint8_t* p) {
  ap_uint<W> v = 0;
  for (int b = 0; b < W; b += 8) {
    int n = (W - b) < 8 ? (W - b) : 8;
    v(b + n - 1, b) = (unsigned long long)(p[b / 8] & ((1u << n) - 1u));
  }
  return v;
}
template <int W> void store_word(uint8_t* p, const ap_uint<W>& v) {
  std::memset(p, 0, container_bytes(W));
  for (int b = 0; b < W; b += 8) {
    int n = (W - b) < 8 ? (W - b) : 8;
    p[b / 8] = (uint8_t)(unsigned long long)v(b + n - 1, b);
  }
}
// signed scalar (thresholds): low W bits of the container, sign-extended by the ap_int ctor
template <int W> ap_int<W> load_sword(const uint8_t* p) {
  ap_uint<W> u = load_word<W>(p);
  ap_int<W> s;
  s(W - 1, 0) = u(W - 1, 0);
  return s;
}

template <int W> void fill_stream(hls::stream<ap_uint<W> >& s, const uint8_t* bytes, size_t nwords) {
  s.reserve(nwords);
  for (size_t i = 0; i < nwords; i++) s.write(load_word<W>(bytes + i * container_bytes(W)));
}
template <int W> int drain_stream(hls::stream<ap_uint<W> >& s, uint8_t* bytes, size_t nwords) {
  if (s.size() != nwords) return -2;
  for (size_t i = 0; i < nwords; i++) store_word<W>(bytes + i * container_bytes(W), s.read());
  return 0;
}

template <unsigned SIMD, typename WT, unsigned PE, unsigned TILES>
void load_weights(FixedPointWeights<SIMD, WT, PE, TILES>& w, const uint8_t* bytes) {
  const int cb = container_bytes(SIMD * WT::width);
  for (unsigned pe = 0; pe < PE; pe++)
    for (unsigned t = 0; t < TILES; t++) w.m_weights[pe][t] = load_word<SIMD * WT::width>(bytes + (size_t)(pe * TILES + t) * cb);
}
template <unsigned SIMD, unsigned PE, unsigned TILES>
void load_bweights(BinaryWeights<SIMD, PE, TILES>& w, const uint8_t* bytes) {
  const int cb = container_bytes(SIMD);
  for (unsigned pe = 0; pe < PE; pe++)
    for (unsigned t = 0; t < TILES; t++) w.m_weights[pe][t] = load_word<SIMD>(bytes + (size_t)(pe * TILES + t) * cb);
}
template <unsigned NF, unsigned PE, unsigned NTH, int TAB, typename TR, int AV>
void load_thresholds(ThresholdsActivation<NF, PE, NTH, ap_int<TAB>, TR, AV>& a, const uint8_t* bytes) {
  const int cb = container_bytes(TAB);
  for (unsigned pe = 0; pe < PE; pe++)
    for (unsigned nf = 0; nf < NF; nf++)
      for (unsigned i = 0; i < NTH; i++) a.m_thresholds[pe][nf][i] = load_sword<TAB>(bytes + (size_t)((pe * NF + nf) * NTH + i) * cb);
}

// ---- conv2d<> (conv_nonsquare_top.cpp:198-280) ---------------------------------
template <unsigned KX, unsigned KY, unsigned SIMD, unsigned PE, unsigned WB, unsigned C, unsigned OFM, unsigned IX, unsigned IY,
          unsigned OX, unsigned OY, unsigned SX, unsigned SY, unsigned PAD, unsigned INB, unsigned ACTB>
int run_conv2d(const uint8_t* in, const uint8_t* wts, const uint8_t* bias, uint8_t* out, double* secs) {
  constexpr unsigned TILES = (KX * KY * C / SIMD) * (OFM / PE);
  static FixedPointWeights<SIMD, ap_int<WB>, PE, TILES> w;
  static FixedPointWeights<1, ap_int<8>, 1, OFM> b;
  load_weights(w, wts);
  load_weights(b, bias);
  hls::stream<ap_uint<C * INB> > s_in("in");
  hls::stream<ap_uint<OFM * ACTB> > s_out("out");
  fill_stream<C * INB>(s_in, in, (size_t)IX * IY);
  auto t0 = std::chrono::steady_clock::now();
  conv2d<KX, KY, SIMD, PE, WB, C, OFM, IX, IY, OX, OY, SX, SY, PAD, INB, TILES, ACTB>(w, b, s_in, s_out, 1);
  auto t1 = std::chrono::steady_clock::now();
  if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
  return drain_stream<OFM * ACTB>(s_out, out, (size_t)OX * OY);
}

// ---- deconv522<> (conv_nonsquare_top.cpp:71-195) -------------------------------
template <unsigned IX, unsigned IY, unsigned C, unsigned OFM, unsigned SIMD, unsigned PE, unsigned WB>
int run_deconv(const uint8_t* in, const uint8_t* wts, const uint8_t* bias, uint8_t* out, double* secs) {
  constexpr unsigned TILES = (25 * C / SIMD) * (OFM / PE);
  static FixedPointWeights<SIMD, ap_int<WB>, PE, TILES> w;
  static FixedPointWeights<1, ap_int<8>, 1, OFM> b;
  load_weights(w, wts);
  load_weights(b, bias);
  hls::stream<ap_uint<C * 8> > s_in("in");
  hls::stream<ap_uint<OFM * 8> > s_out("out");
  fill_stream<C * 8>(s_in, in, (size_t)IX * IY);
  auto t0 = std::chrono::steady_clock::now();
  deconv522<IX, IY, C, OFM, SIMD, PE, 8, 8, WB, TILES>(w, b, s_in, s_out, 1);
  auto t1 = std::chrono::steady_clock::now();
  if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
  return drain_stream<OFM * 8>(s_out, out, (size_t)(2 * IX) * (2 * IY));
}

// ---- library path, as the commented Testbench_conv_nonsquare composes it
//      (conv_nonsquare_top.cpp:361-379), plus FMPadding_nonsquare in front
//      (streamtools.h:361-406) when PAD > 0: fixed-point weights + ThresholdsActivation.
template <unsigned K, unsigned SIMD, unsigned PE, unsigned WB, unsigned C, unsigned OFM, unsigned IX, unsigned IY, unsigned PAD,
          unsigned INB, unsigned NTH, int TAB, unsigned TRB, int AV>
int run_thresh(const uint8_t* in, const uint8_t* wts, const uint8_t* thr, uint8_t* out, double* secs) {
  constexpr unsigned PX = IX + 2 * PAD, PY = IY + 2 * PAD, OX = PX - K + 1, OY = PY - K + 1;
  constexpr unsigned MW = K * K * C, MH = OFM, SF = MW / SIMD, NF = MH / PE;
  static FixedPointWeights<SIMD, ap_int<WB>, PE, SF * NF> w;
  static ThresholdsActivation<NF, PE, NTH, ap_int<TAB>, ap_uint<TRB>, AV> act;
  load_weights(w, wts);
  load_thresholds(act, thr);
  hls::stream<ap_uint<C * INB> > s_in("in"), s_pad("pad");
  hls::stream<ap_uint<SIMD * INB> > s_wa("wa"), s_win("win");
  hls::stream<ap_uint<PE * TRB> > s_mv("mv");
  hls::stream<ap_uint<OFM * TRB> > s_out("out");
  fill_stream<C * INB>(s_in, in, (size_t)IX * IY);
  auto t0 = std::chrono::steady_clock::now();
  FMPadding_nonsquare<PX, PY, 2 * PAD, 2 * PAD, C, C, ap_uint<INB> >(s_in, s_pad);
  StreamingDataWidthConverter_Batch<C * INB, SIMD * INB, PX * PY>(s_pad, s_wa, 1);
  ConvolutionInputGenerator_NonSquare<K, K, C, INB, PX, PY, OX, OY, SIMD, 1, 1>(s_wa, s_win, 1, ap_resource_dflt());
  Matrix_Vector_Activate_Batch<MW, MH, SIMD, PE, 1, Slice<ap_uint<INB> >, Slice<ap_uint<TRB> >, Identity>(
      s_win, s_mv, w, act, OX * OY, ap_resource_dsp());
  StreamingDataWidthConverter_Batch<PE * TRB, OFM * TRB, OX * OY * NF>(s_mv, s_out, 1);
  auto t1 = std::chrono::steady_clock::now();
  if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
  return drain_stream<OFM * TRB>(s_out, out, (size_t)OX * OY);
}

// ---- the same library path with the weights arriving as a STREAM: GenParamStream (dma.h:214-236) feeding
//      Matrix_Vector_Activate_Stream_Batch (mvau.hpp:209-307).  Also hands back the first period (TILES words) of the
//      parameter stream, the image fcb_layer_set_param_stream takes.
template <unsigned K, unsigned SIMD, unsigned PE, unsigned WB, unsigned C, unsigned OFM, unsigned IX, unsigned IY, unsigned PAD,
          unsigned INB, unsigned NTH, int TAB, unsigned TRB, int AV>
int run_thresh_stream(const uint8_t* in, const uint8_t* wts, const uint8_t* thr, uint8_t* out, uint8_t* param_words) {
  constexpr unsigned PX = IX + 2 * PAD, PY = IY + 2 * PAD, OX = PX - K + 1, OY = PY - K + 1;
  constexpr unsigned MW = K * K * C, MH = OFM, SF = MW / SIMD, NF = MH / PE;
  static FixedPointWeights<SIMD, ap_int<WB>, PE, SF * NF> w;
  static ThresholdsActivation<NF, PE, NTH, ap_int<TAB>, ap_uint<TRB>, AV> act;
  load_weights(w, wts);
  load_thresholds(act, thr);
  hls::stream<ap_uint<C * INB> > s_in("in"), s_pad("pad");
  hls::stream<ap_uint<SIMD * INB> > s_wa("wa"), s_win("win");
  hls::stream<ap_uint<SIMD * PE * WB> > s_par("par"), s_one("one");
  hls::stream<ap_uint<PE * TRB> > s_mv("mv");
  hls::stream<ap_uint<OFM * TRB> > s_out("out");
  GenParamStream<SF * NF, SIMD, PE, WB>(w, s_one, 1);
  if (drain_stream<SIMD * PE * WB>(s_one, param_words, (size_t)SF * NF)) return -3;
  fill_stream<C * INB>(s_in, in, (size_t)IX * IY);
  FMPadding_nonsquare<PX, PY, 2 * PAD, 2 * PAD, C, C, ap_uint<INB> >(s_in, s_pad);
  StreamingDataWidthConverter_Batch<C * INB, SIMD * INB, PX * PY>(s_pad, s_wa, 1);
  ConvolutionInputGenerator_NonSquare<K, K, C, INB, PX, PY, OX, OY, SIMD, 1, 1>(s_wa, s_win, 1, ap_resource_dflt());
  GenParamStream<SF * NF, SIMD, PE, WB>(w, s_par, OX * OY);
  Matrix_Vector_Activate_Stream_Batch<MW, MH, SIMD, PE, Slice<ap_uint<INB> >, Slice<ap_uint<TRB> >, Identity, ap_int<WB> >(
      s_win, s_mv, s_par, act, OX * OY, ap_resource_dsp());
  StreamingDataWidthConverter_Batch<PE * TRB, OFM * TRB, OX * OY * NF>(s_mv, s_out, 1);
  return drain_stream<OFM * TRB>(s_out, out, (size_t)OX * OY);
}

// ---- 1-bit path: BinaryWeights + Recast<XnorMul> + ThresholdsActivation (NumTH=1, TR=ap_uint<1>)
//      weights.hpp:66-98, interpret.hpp:57-73,126-175, activations.hpp:168-190; no padding (ConvLayer_Batch semantics)
template <unsigned K, unsigned SIMD, unsigned PE, unsigned C, unsigned OFM, unsigned IX, unsigned IY, int TAB>
int run_xnor(const uint8_t* in, const uint8_t* wts, const uint8_t* thr, uint8_t* out, double* secs) {
  constexpr unsigned OX = IX - K + 1, OY = IY - K + 1;
  constexpr unsigned MW = K * K * C, MH = OFM, SF = MW / SIMD, NF = MH / PE;
  static BinaryWeights<SIMD, PE, SF * NF> w;
  static ThresholdsActivation<NF, PE, 1, ap_int<TAB>, ap_uint<1> > act;
  load_bweights(w, wts);
  load_thresholds(act, thr);
  hls::stream<ap_uint<C> > s_in("in");
  hls::stream<ap_uint<SIMD> > s_wa("wa"), s_win("win");
  hls::stream<ap_uint<PE> > s_mv("mv");
  hls::stream<ap_uint<OFM> > s_out("out");
  fill_stream<C>(s_in, in, (size_t)IX * IY);
  auto t0 = std::chrono::steady_clock::now();
  StreamingDataWidthConverter_Batch<C, SIMD, IX * IY>(s_in, s_wa, 1);
  ConvolutionInputGenerator_NonSquare<K, K, C, 1, IX, IY, OX, OY, SIMD, 1, 1>(s_wa, s_win, 1, ap_resource_dflt());
  Matrix_Vector_Activate_Batch<MW, MH, SIMD, PE, 1, Recast<XnorMul>, Slice<ap_uint<1> >, Identity>(
      s_win, s_mv, w, act, OX * OY, ap_resource_lut());
  StreamingDataWidthConverter_Batch<PE, OFM, OX * OY * NF>(s_mv, s_out, 1);
  auto t1 = std::chrono::steady_clock::now();
  if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
  return drain_stream<OFM>(s_out, out, (size_t)OX * OY);
}

// ---- pooling (maxpool.h:66-96, 137-185); the reference is square-only
template <unsigned DIM, unsigned PD, unsigned C, unsigned B>
int run_pool_prec(const uint8_t* in, uint8_t* out) {
  hls::stream<ap_uint<C * B> > s_in("in"), s_out("out");
  fill_stream<C * B>(s_in, in, (size_t)DIM * DIM);
  StreamingMaxPool_Precision<DIM, PD, C, ap_uint<B>, 0>(s_in, s_out);
  return drain_stream<C * B>(s_out, out, (size_t)(DIM / PD) * (DIM / PD));
}
template <unsigned DIM, unsigned PD, unsigned C>
int run_pool_bin(const uint8_t* in, uint8_t* out) {
  hls::stream<ap_uint<C> > s_in("in"), s_out("out");
  fill_stream<C>(s_in, in, (size_t)DIM * DIM);
  StreamingMaxPool<DIM, PD, C>(s_in, s_out);
  return drain_stream<C>(s_out, out, (size_t)(DIM / PD) * (DIM / PD));
}

// ---- the threshold path with FMPadding_nonsquare's own parameters (odd / asymmetric totals, PaddingStyle; streamtools.h:361-379)
//      and optionally StreamingMaxPool_Precision<OX, PD, OFM, ActType, MINV> behind it (maxpool.h:137-185; square maps only)
template <unsigned K, unsigned SIMD, unsigned PE, unsigned WB, unsigned C, unsigned OFM, unsigned IX, unsigned IY, unsigned PXT, unsigned PYT,
          unsigned STYLE, unsigned INB, unsigned NTH, int TAB, unsigned TRB, int AV, unsigned PD, typename POOLT, int MINV>
int run_thresh_pad_pool(const uint8_t* in, const uint8_t* wts, const uint8_t* thr, uint8_t* out) {
  constexpr unsigned PX = IX + PXT, PY = IY + PYT, OX = PX - K + 1, OY = PY - K + 1;
  constexpr unsigned MW = K * K * C, MH = OFM, SF = MW / SIMD, NF = MH / PE;
  static FixedPointWeights<SIMD, ap_int<WB>, PE, SF * NF> w;
  static ThresholdsActivation<NF, PE, NTH, ap_int<TAB>, ap_uint<TRB>, AV> act;
  load_weights(w, wts);
  load_thresholds(act, thr);
  hls::stream<ap_uint<C * INB> > s_in("in"), s_pad("pad");
  hls::stream<ap_uint<SIMD * INB> > s_wa("wa"), s_win("win");
  hls::stream<ap_uint<PE * TRB> > s_mv("mv");
  hls::stream<ap_uint<OFM * TRB> > s_act("act"), s_out("out");
  fill_stream<C * INB>(s_in, in, (size_t)IX * IY);
  FMPadding_nonsquare<PX, PY, PXT, PYT, C, C, ap_uint<INB>, STYLE>(s_in, s_pad);
  StreamingDataWidthConverter_Batch<C * INB, SIMD * INB, PX * PY>(s_pad, s_wa, 1);
  ConvolutionInputGenerator_NonSquare<K, K, C, INB, PX, PY, OX, OY, SIMD, 1, 1>(s_wa, s_win, 1, ap_resource_dflt());
  Matrix_Vector_Activate_Batch<MW, MH, SIMD, PE, 1, Slice<ap_uint<INB> >, Slice<ap_uint<TRB> >, Identity>(
      s_win, s_mv, w, act, OX * OY, ap_resource_dsp());
  StreamingDataWidthConverter_Batch<PE * TRB, OFM * TRB, OX * OY * NF>(s_mv, s_act, 1);
  if (PD <= 1) return drain_stream<OFM * TRB>(s_act, out, (size_t)OX * OY);
  static_assert(PD <= 1 || OX == OY, "StreamingMaxPool_Precision is square-only");
  StreamingMaxPool_Precision<OX, (PD > 1 ? PD : 1), OFM, POOLT, MINV>(s_act, s_out);
  return drain_stream<OFM * TRB>(s_out, out, (size_t)(OX / (PD > 1 ? PD : 1)) * (OY / (PD > 1 ? PD : 1)));
}

// ---- wide types: signed 16-bit lanes, 16-bit weights, PassThroughActivation<ap_int<TAB>> with TAB > 32 (mvau.hpp:112 allows any TA),
//      32-bit output lanes (the range assignment at mvau.hpp:167 truncates)
template <unsigned K, unsigned SIMD, unsigned PE, unsigned WB, unsigned C, unsigned OFM, unsigned IX, unsigned IY, unsigned PAD, unsigned INB,
          int TAB, unsigned OUTB>
int run_wide(const uint8_t* in, const uint8_t* wts, uint8_t* out) {
  constexpr unsigned PX = IX + 2 * PAD, PY = IY + 2 * PAD, OX = PX - K + 1, OY = PY - K + 1;
  constexpr unsigned MW = K * K * C, MH = OFM, SF = MW / SIMD, NF = MH / PE;
  static FixedPointWeights<SIMD, ap_int<WB>, PE, SF * NF> w;
  load_weights(w, wts);
  hls::stream<ap_uint<C * INB> > s_in("in"), s_pad("pad");
  hls::stream<ap_uint<SIMD * INB> > s_wa("wa"), s_win("win");
  hls::stream<ap_uint<PE * OUTB> > s_mv("mv");
  hls::stream<ap_uint<OFM * OUTB> > s_out("out");
  fill_stream<C * INB>(s_in, in, (size_t)IX * IY);
  FMPadding_nonsquare<PX, PY, 2 * PAD, 2 * PAD, C, C, ap_uint<INB> >(s_in, s_pad);
  StreamingDataWidthConverter_Batch<C * INB, SIMD * INB, PX * PY>(s_pad, s_wa, 1);
  ConvolutionInputGenerator_NonSquare<K, K, C, INB, PX, PY, OX, OY, SIMD, 1, 1>(s_wa, s_win, 1, ap_resource_dflt());
  Matrix_Vector_Activate_Batch<MW, MH, SIMD, PE, 1, Slice<ap_int<INB> >, Slice<ap_uint<OUTB> >, Identity>(
      s_win, s_mv, w, PassThroughActivation<ap_int<TAB> >(), OX * OY, ap_resource_dsp());
  StreamingDataWidthConverter_Batch<PE * OUTB, OFM * OUTB, OX * OY * NF>(s_mv, s_out, 1);
  return drain_stream<OFM * OUTB>(s_out, out, (size_t)OX * OY);
}

// ---- ConvolutionInputGenerator_NonSquare_Dilated (slidingwindow.h:1515-1631; x dilation only) and
//      ConvolutionInputGenerator_kernel_stride (K % S != 0, square; :447-575) in front of the MVAU with thresholds
template <unsigned KX, unsigned KY, unsigned SIMD, unsigned PE, unsigned WB, unsigned C, unsigned OFM, unsigned IX, unsigned IY, unsigned S, unsigned DX,
          unsigned KSTRIDE_GEN, unsigned NTH, int TAB, unsigned TRB>
int run_swg_variant(const uint8_t* in, const uint8_t* wts, const uint8_t* thr, uint8_t* out) {
  constexpr unsigned OX = (IX - ((KX - 1) * DX + 1)) / S + 1, OY = (IY - KY) / S + 1;
  constexpr unsigned MW = KX * KY * C, MH = OFM, SF = MW / SIMD, NF = MH / PE;
  static FixedPointWeights<SIMD, ap_int<WB>, PE, SF * NF> w;
  static ThresholdsActivation<NF, PE, NTH, ap_int<TAB>, ap_uint<TRB>, 0> act;
  load_weights(w, wts);
  load_thresholds(act, thr);
  hls::stream<ap_uint<C * 8> > s_in("in");
  hls::stream<ap_uint<SIMD * 8> > s_wa("wa"), s_win("win");
  hls::stream<ap_uint<PE * TRB> > s_mv("mv");
  hls::stream<ap_uint<OFM * TRB> > s_out("out");
  fill_stream<C * 8>(s_in, in, (size_t)IX * IY);
  StreamingDataWidthConverter_Batch<C * 8, SIMD * 8, IX * IY>(s_in, s_wa, 1);
  if (KSTRIDE_GEN) ConvolutionInputGenerator_kernel_stride<KX, C, 8, IX, OX, SIMD, S>(s_wa, s_win, 1, ap_resource_dflt());
  else ConvolutionInputGenerator_NonSquare_Dilated<KX, KY, C, 8, IX, IY, OX, OY, SIMD, S, S, DX, 1>(s_wa, s_win, 1, ap_resource_dflt());
  Matrix_Vector_Activate_Batch<MW, MH, SIMD, PE, 1, Slice<ap_uint<8> >, Slice<ap_uint<TRB> >, Identity>(s_win, s_mv, w, act, OX * OY,
                                                                                                     ap_resource_dsp());
  StreamingDataWidthConverter_Batch<PE * TRB, OFM * TRB, OX * OY * NF>(s_mv, s_out, 1);
  return drain_stream<OFM * TRB>(s_out, out, (size_t)OX * OY);
}

// ---- channel-wise units behind the depth-wise sliding window: ConvolutionInputGenerator_dws (square, any stride with K % S == 0,
//      slidingwindow.h:761-868) or ConvolutionInputGenerator_NonSquare_dws (stride 1, :1377-1488), FMPadding_nonsquare in front.
template <unsigned K, unsigned C, unsigned PE, unsigned IX, unsigned IY, unsigned S, unsigned PAD, unsigned INB>
void dws_window(hls::stream<ap_uint<C * INB> >& s_in, hls::stream<ap_uint<PE * INB> >& s_win) {
  constexpr unsigned PX = IX + 2 * PAD, PY = IY + 2 * PAD, OX = (PX - K) / S + 1, OY = (PY - K) / S + 1;
  hls::stream<ap_uint<C * INB> > s_pad("pad");
  hls::stream<ap_uint<PE * INB> > s_wa("wa");
  FMPadding_nonsquare<PX, PY, 2 * PAD, 2 * PAD, C, C, ap_uint<INB> >(s_in, s_pad);
  StreamingDataWidthConverter_Batch<C * INB, PE * INB, PX * PY>(s_pad, s_wa, 1);
  if (PX == PY) ConvolutionInputGenerator_dws<K, C, INB, PX, OX, PE, S>(s_wa, s_win, 1, ap_resource_dflt());
  else ConvolutionInputGenerator_NonSquare_dws<K, K, C, INB, PX, PY, OX, OY, PE, 1, 1>(s_wa, s_win, 1, ap_resource_dflt());
}

// Pool_batch (maxpool.h:525-577) with a pool.hpp function object
template <unsigned K, unsigned C, unsigned PE, unsigned IX, unsigned IY, unsigned S, unsigned PAD, unsigned INB, typename TIN, unsigned OUTB,
          typename FN>
int run_pool_batch(const uint8_t* in, uint8_t* out) {
  constexpr unsigned PX = IX + 2 * PAD, PY = IY + 2 * PAD, OX = (PX - K) / S + 1, OY = (PY - K) / S + 1;
  static_assert(PX == PY || S == 1, "the non-square depth-wise generator is used at stride 1 only");
  hls::stream<ap_uint<C * INB> > s_in("in");
  hls::stream<ap_uint<PE * INB> > s_win("win");
  hls::stream<ap_uint<PE * OUTB> > s_p("p");
  hls::stream<ap_uint<C * OUTB> > s_out("out");
  fill_stream<C * INB>(s_in, in, (size_t)IX * IY);
  dws_window<K, C, PE, IX, IY, S, PAD, INB>(s_in, s_win);
  Pool_batch<C, PE, K, Slice<TIN>, Slice<ap_uint<OUTB> > >(s_win, s_p, FN(), OX * OY);
  StreamingDataWidthConverter_Batch<PE * OUTB, C * OUTB, OX * OY * (C / PE)>(s_p, s_out, 1);
  return drain_stream<C * OUTB>(s_out, out, (size_t)OX * OY);
}

// Vector_Vector_Activate_Batch (vvau.hpp:80-154): FixedPointWeights<1, ap_int<WB>, PE, NF*K*K>, any activation object
template <unsigned K, unsigned C, unsigned PE, unsigned IX, unsigned IY, unsigned S, unsigned PAD, unsigned INB, unsigned WB, unsigned OUTB,
          typename ACT>
int run_vvau(const uint8_t* in, const uint8_t* wts, ACT& act, uint8_t* out) {
  constexpr unsigned PX = IX + 2 * PAD, PY = IY + 2 * PAD, OX = (PX - K) / S + 1, OY = (PY - K) / S + 1, NF = C / PE;
  static_assert(PX == PY || S == 1, "the non-square depth-wise generator is used at stride 1 only");
  static FixedPointWeights<1, ap_int<WB>, PE, NF * K * K> w;
  load_weights(w, wts);
  hls::stream<ap_uint<C * INB> > s_in("in");
  hls::stream<ap_uint<PE * INB> > s_win("win");
  hls::stream<ap_uint<PE * OUTB> > s_p("p");
  hls::stream<ap_uint<C * OUTB> > s_out("out");
  fill_stream<C * INB>(s_in, in, (size_t)IX * IY);
  dws_window<K, C, PE, IX, IY, S, PAD, INB>(s_in, s_win);
  Vector_Vector_Activate_Batch<C, K * K, PE, PE, 1, Slice<ap_uint<INB> >, Slice<ap_uint<OUTB> >, Identity>(s_win, s_p, w, act, OX * OY,
                                                                                                          ap_resource_dsp());
  StreamingDataWidthConverter_Batch<PE * OUTB, C * OUTB, OX * OY * NF>(s_p, s_out, 1);
  return drain_stream<C * OUTB>(s_out, out, (size_t)OX * OY);
}

}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// padding / pool / lane-width forms of the threshold path: ref_<name>(in, weights, thresholds, out, secs)
//   K SIMD PE WB  C OFM IX IY  PadX PadY Style INB NTH TAB TRB AV  PoolDim ActType MinValue
REF_API int ref_px_odd2(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // 3 left/2... : style 2 puts the odd zero left / up
  return run_thresh_pad_pool<3, 4, 2, 4, 8, 8, 10, 6, 3, 1, 2, 8, 15, 24, 4, 0, 1, ap_uint<4>, 0>(in, w, t, out);
}
REF_API int ref_px_odd1(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // other styles: the odd zero goes right / down
  return run_thresh_pad_pool<3, 4, 2, 4, 8, 8, 10, 6, 3, 1, 1, 8, 15, 24, 4, 0, 1, ap_uint<4>, 0>(in, w, t, out);
}
REF_API int ref_pk3_signed(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // 12x12 map, PoolDim 3, ap_int<4> compare from -8
  return run_thresh_pad_pool<3, 4, 2, 4, 8, 8, 12, 12, 2, 2, 2, 8, 15, 24, 4, 0, 3, ap_int<4>, -8>(in, w, t, out);
}
REF_API int ref_pk2_min5(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // PoolDim 2, maxima start from 5
  return run_thresh_pad_pool<3, 4, 2, 4, 8, 8, 12, 12, 2, 2, 2, 8, 15, 24, 4, 0, 2, ap_uint<4>, 5>(in, w, t, out);
}
REF_API int ref_lw3(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // ap_uint<3> activations, ap_int<3> weights, ap_uint<3> outputs
  return run_thresh_pad_pool<3, 4, 2, 3, 8, 8, 10, 6, 2, 2, 2, 3, 7, 12, 3, 0, 1, ap_uint<3>, 0>(in, w, t, out);
}
REF_API int ref_lw5x12(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // ap_uint<5> in, ap_int<5> weights, 12 channels of ap_uint<6> out
  return run_thresh_pad_pool<3, 3, 4, 5, 6, 12, 9, 7, 0, 0, 2, 5, 40, 16, 6, 0, 1, ap_uint<6>, 0>(in, w, t, out);
}

REF_API int ref_acc40(const uint8_t* in, const uint8_t* w, const uint8_t*, uint8_t* out, double*) {  // s16 x s16 -> ap_int<40> -> 32-bit lanes
  return run_wide<3, 4, 2, 16, 8, 8, 10, 6, 1, 16, 40, 32>(in, w, out);
}

REF_API int ref_dil_x2(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // 3x3, Dilation_x 2, 14x8 -> 10x6
  return run_swg_variant<3, 3, 4, 2, 4, 8, 8, 14, 8, 1, 2, 0, 15, 24, 4>(in, w, t, out);
}
REF_API int ref_dil_x3_k2(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // 2x3 kernel (Kx 2, Ky 3), Dilation_x 3, 16 ch
  return run_swg_variant<2, 3, 8, 4, 4, 16, 8, 13, 7, 1, 3, 0, 15, 24, 4>(in, w, t, out);
}
REF_API int ref_ks_k3s2(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {  // kernel_stride generator: K 3, S 2 (3 % 2 != 0), 11x11 -> 5x5
  return run_swg_variant<3, 3, 4, 2, 4, 8, 8, 11, 11, 2, 1, 1, 15, 24, 4>(in, w, t, out);
}

// AddStreams_Batch (streamtools.h:669-720): ref_<name>(in1, in2, out)
template <unsigned CH, typename T1, typename T2, typename TO, unsigned NTOT, int OFF>
int run_add(const uint8_t* in1, const uint8_t* in2, uint8_t* out, unsigned reps) {
  hls::stream<ap_uint<CH * T1::width> > s1("s1");
  hls::stream<ap_uint<CH * T2::width> > s2("s2");
  hls::stream<ap_uint<CH * TO::width> > so("so");
  fill_stream<CH * T1::width>(s1, in1, (size_t)NTOT * reps);
  fill_stream<CH * T2::width>(s2, in2, (size_t)NTOT * reps);
  AddStreams_Batch<CH, T1, T2, TO, NTOT, OFF>(s1, s2, so, reps);
  return drain_stream<CH * TO::width>(so, out, (size_t)NTOT * reps);
}
REF_API int ref_add_u8(const uint8_t* a, const uint8_t* b, uint8_t* o) { return run_add<16, ap_uint<8>, ap_uint<8>, ap_uint<8>, 40, 0>(a, b, o, 3); }
REF_API int ref_add_s8_off(const uint8_t* a, const uint8_t* b, uint8_t* o) { return run_add<6, ap_int<8>, ap_uint<4>, ap_int<10>, 33, -7>(a, b, o, 2); }

// StreamingFCLayer_Batch (fclayer.h:83-111): ref_fc_<name>(in, weights, thresholds-or-unused, out, secs); one input word of MatrixW
// lanes and one output word of MatrixH lanes per repetition
template <unsigned MW, unsigned MH, unsigned SIMD, unsigned PE, unsigned INB, unsigned WB, int TAB, unsigned OUTB, unsigned REPS>
int run_fc_pass(const uint8_t* in, const uint8_t* wts, uint8_t* out) {
  static FixedPointWeights<SIMD, ap_int<WB>, PE, (MW / SIMD) * (MH / PE)> w;
  load_weights(w, wts);
  hls::stream<ap_uint<MW * INB> > s_in("in");
  hls::stream<ap_uint<MH * OUTB> > s_out("out");
  fill_stream<MW * INB>(s_in, in, REPS);
  StreamingFCLayer_Batch<MW, MH, SIMD, PE, Slice<ap_uint<INB> >, Slice<ap_uint<OUTB> >, Identity>(s_in, s_out, w, PassThroughActivation<ap_int<TAB> >(),
                                                                                                  REPS, ap_resource_dsp());
  return drain_stream<MH * OUTB>(s_out, out, REPS);
}
REF_API int ref_fc_a(const uint8_t* in, const uint8_t* w, const uint8_t*, uint8_t* out, double*) {
  return run_fc_pass<64, 32, 8, 4, 8, 4, 16, 16, 7>(in, w, out);
}

// Pool_batch cases: ref_<name>(in, unused, unused, out, secs)
//                         K  C  PE IX  IY  S PAD INB  TSrcI lane     OUTB  function
REF_API int ref_pl_max_a(const uint8_t* in, const uint8_t*, const uint8_t*, uint8_t* out, double*) {
  return run_pool_batch<2, 8, 4, 12, 12, 2, 0, 8, ap_uint<8>, 8, MaxPoolFunction<ap_uint<8>, 2> >(in, out);
}
REF_API int ref_pl_max_s(const uint8_t* in, const uint8_t*, const uint8_t*, uint8_t* out, double*) {
  return run_pool_batch<3, 4, 2, 10, 6, 1, 1, 8, ap_int<8>, 8, MaxPoolFunction<ap_int<8>, 3> >(in, out);
}
REF_API int ref_pl_avg(const uint8_t* in, const uint8_t*, const uint8_t*, uint8_t* out, double*) {
  return run_pool_batch<2, 8, 8, 8, 8, 2, 0, 8, ap_uint<8>, 8, AvgPoolFunction<ap_uint<10>, ap_uint<8>, 4> >(in, out);
}
REF_API int ref_pl_qavg(const uint8_t* in, const uint8_t*, const uint8_t*, uint8_t* out, double*) {
  return run_pool_batch<4, 4, 4, 16, 16, 4, 0, 8, ap_int<8>, 8, QuantAvgPoolFunction<ap_int<12>, ap_int<8>, 4> >(in, out);
}
REF_API int ref_pl_acc(const uint8_t* in, const uint8_t*, const uint8_t*, uint8_t* out, double*) {
  return run_pool_batch<3, 6, 3, 9, 7, 1, 0, 4, ap_uint<4>, 8, AccPoolFunction<ap_uint<8>, 9> >(in, out);
}
// depth-wise convolution cases: ref_<name>(in, weights, thresholds-or-unused, out, secs)
REF_API int ref_dw_a(const uint8_t* in, const uint8_t* w, const uint8_t*, uint8_t* out, double*) {
  PassThroughActivation<ap_int<16> > act;
  return run_vvau<3, 8, 4, 10, 6, 1, 1, 8, 4, 16>(in, w, act, out);
}
REF_API int ref_dw_b(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double*) {
  static ThresholdsActivation<2, 8, 15, ap_int<16>, ap_uint<4>, 0> act;
  load_thresholds(act, t);
  return run_vvau<3, 16, 8, 12, 12, 1, 1, 8, 4, 4>(in, w, act, out);
}
REF_API int ref_dw_c(const uint8_t* in, const uint8_t* w, const uint8_t*, uint8_t* out, double*) {
  PassThroughActivation<ap_int<12> > act;
  return run_vvau<2, 4, 2, 8, 8, 2, 0, 8, 4, 12>(in, w, act, out);
}

// conv2d<> instantiations.  Name, KX,KY, SIMD,PE, WB, C,OFM, IX,IY, OX,OY, S, PAD, INB, ACTB
#define CONV2D_CASES(X)                                                   \
  X(c2d_a, 5, 5, 2, 3, 4, 4, 6, 12, 8, 6, 4, 2, 2, 8, 8)                  \
  X(c2d_b, 5, 5, 8, 8, 4, 16, 32, 40, 24, 20, 12, 2, 2, 8, 8)             \
  X(c2d_c, 5, 5, 3, 8, 4, 3, 16, 32, 20, 16, 10, 2, 2, 8, 8)              \
  X(c2d_d, 3, 3, 16, 4, 8, 32, 32, 20, 12, 20, 12, 1, 1, 8, 8)            \
  X(c2d_e, 5, 5, 8, 16, 4, 128, 128, 48, 32, 24, 16, 2, 2, 8, 8)          \
  X(c2d_f, 3, 3, 4, 2, 4, 8, 8, 10, 6, 10, 6, 1, 1, 8, 16)                \
  X(c2d_g, 5, 5, 8, 24, 4, 128, 192, 24, 16, 12, 8, 2, 2, 8, 8)           \
  X(c2d_L1band, 5, 5, 8, 16, 4, 128, 128, 384, 32, 192, 16, 2, 2, 8, 8)   \
  X(c2d_L1, 5, 5, 8, 16, 4, 128, 128, 384, 256, 192, 128, 2, 2, 8, 8)

#define X(name, KX, KY, SIMD, PE, WB, C, OFM, IX, IY, OX, OY, S, PAD, INB, ACTB)                               \
  REF_API int ref_##name(const uint8_t* in, const uint8_t* w, const uint8_t* b, uint8_t* out, double* secs) {  \
    return run_conv2d<KX, KY, SIMD, PE, WB, C, OFM, IX, IY, OX, OY, S, S, PAD, INB, ACTB>(in, w, b, out, secs); \
  }
CONV2D_CASES(X)
#undef X

// deconv522<> instantiations.  Name, IX,IY, C,OFM, SIMD,PE, WB
#define DECONV_CASES(X)                        \
  X(dc_a, 6, 4, 4, 6, 2, 3, 4)                 \
  X(dc_b, 12, 8, 16, 16, 8, 8, 4)              \
  X(dc_c, 24, 16, 128, 128, 8, 16, 4)          \
  X(dc_d, 24, 16, 128, 3, 8, 3, 4)             \
  X(dc_e, 12, 8, 192, 128, 12, 16, 4)          \
  X(dc_L4, 48, 32, 192, 128, 12, 16, 4)

#define X(name, IX, IY, C, OFM, SIMD, PE, WB)                                                                  \
  REF_API int ref_##name(const uint8_t* in, const uint8_t* w, const uint8_t* b, uint8_t* out, double* secs) {  \
    return run_deconv<IX, IY, C, OFM, SIMD, PE, WB>(in, w, b, out, secs);                                       \
  }
DECONV_CASES(X)
#undef X

// threshold path.  Name, K, SIMD,PE, WB, C,OFM, IX,IY, PAD, INB, NTH, TAB, TRB, ActVal
#define THRESH_CASES(X)                                            \
  X(th_a, 3, 4, 2, 4, 8, 8, 10, 6, 1, 8, 15, 24, 4, 0)             \
  X(th_b, 3, 16, 8, 4, 32, 32, 16, 12, 1, 8, 255, 24, 8, 0)        \
  X(th_c, 3, 8, 4, 4, 16, 16, 9, 7, 0, 8, 3, 16, 2, 0)             \
  X(th_d, 3, 3, 8, 4, 3, 16, 16, 12, 1, 8, 255, 24, 8, 0)          \
  X(th_cfg4, 3, 32, 32, 4, 256, 256, 64, 48, 1, 8, 255, 24, 8, 0)

#define X(name, K, SIMD, PE, WB, C, OFM, IX, IY, PAD, INB, NTH, TAB, TRB, AV)                                   \
  REF_API int ref_##name(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double* secs) {  \
    return run_thresh<K, SIMD, PE, WB, C, OFM, IX, IY, PAD, INB, NTH, TAB, TRB, AV>(in, w, t, out, secs);       \
  }
THRESH_CASES(X)
#undef X

// streamed-weights form of the small threshold cases: ref_stream_<name>(in, weights, thresholds, out, param_words_out)
#define STREAM_CASES(X)                                            \
  X(th_a, 3, 4, 2, 4, 8, 8, 10, 6, 1, 8, 15, 24, 4, 0)             \
  X(th_b, 3, 16, 8, 4, 32, 32, 16, 12, 1, 8, 255, 24, 8, 0)        \
  X(th_c, 3, 8, 4, 4, 16, 16, 9, 7, 0, 8, 3, 16, 2, 0)
#define X(name, K, SIMD, PE, WB, C, OFM, IX, IY, PAD, INB, NTH, TAB, TRB, AV)                                         \
  REF_API int ref_stream_##name(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, uint8_t* pw) { \
    return run_thresh_stream<K, SIMD, PE, WB, C, OFM, IX, IY, PAD, INB, NTH, TAB, TRB, AV>(in, w, t, out, pw);        \
  }
STREAM_CASES(X)
#undef X

// xnor path.  Name, K, SIMD,PE, C,OFM, IX,IY, TAB
#define XNOR_CASES(X)                          \
  X(xn_a, 3, 8, 4, 8, 8, 12, 10, 16)           \
  X(xn_b, 3, 64, 16, 64, 64, 16, 12, 16)       \
  X(xn_c, 3, 32, 8, 64, 32, 20, 9, 16)

#define X(name, K, SIMD, PE, C, OFM, IX, IY, TAB)                                                              \
  REF_API int ref_##name(const uint8_t* in, const uint8_t* w, const uint8_t* t, uint8_t* out, double* secs) {  \
    return run_xnor<K, SIMD, PE, C, OFM, IX, IY, TAB>(in, w, t, out, secs);                                     \
  }
XNOR_CASES(X)
#undef X

REF_API int ref_pool_prec_8_2_8(const uint8_t* in, uint8_t* out) { return run_pool_prec<8, 2, 8, 8>(in, out); }
REF_API int ref_pool_prec_12_2_16(const uint8_t* in, uint8_t* out) { return run_pool_prec<12, 2, 16, 8>(in, out); }
REF_API int ref_pool_prec_12_3_4(const uint8_t* in, uint8_t* out) { return run_pool_prec<12, 3, 4, 4>(in, out); }
REF_API int ref_pool_bin_8_2_16(const uint8_t* in, uint8_t* out) { return run_pool_bin<8, 2, 16>(in, out); }
REF_API int ref_pool_bin_12_2_64(const uint8_t* in, uint8_t* out) { return run_pool_bin<12, 2, 64>(in, out); }

REF_API const char* ref_layers_cases(void) {
  return
#define X(name, ...) #name " "
      CONV2D_CASES(X) DECONV_CASES(X) THRESH_CASES(X) XNOR_CASES(X)
#undef X
      ;
}
