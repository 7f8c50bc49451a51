// oracle/adapter_check.cpp -- TEST INFRASTRUCTURE ONLY (includes reference sources from /root/reference; never part of the product).
//
// The drop-in surface of include/finnconv_hls_adapter.hpp against the REFERENCE's own functions on one small conv2d<> instantiation:
//   1. run_streams over fcb_layer_run                       == conv2d<> (conv_nonsquare_top.cpp:198-280), image by image
//   2. run_streams over fcb_pool_run (replicas on every GPU) == the same
//   3. run_qdma_streams                                     == Qdma2Stream_Batch -> conv2d<> -> Stream2Qdma_Batch (streamtools.h:1001-1037),
//                                                              TKEEP / TLAST included
//   4. run_axi_memory (64-bit memory words)                 == Mem2Stream_Batch -> DWC -> conv2d<> -> DWC -> Stream2Mem_Batch
//                                                              (dma.h:135-199 incl. its 16-image bursts, streamtools.h:463-526)
// Built by `make -C oracle adapter_check` into oracle/_ref/adapter_check where /root/reference exists; run by
// tests/test_bench_scale.py::test_adapter_against_reference_functions on the GPU box.  Prints "adapter_check: all N checks passed".
#define AP_INT_MAX_W 16384
#include <hls_stream.h>
#include "ap_int.h"
#include "ap_axi_sdata.h"
#include "weights.hpp"
#include "config_nonsquare.h"

#define PARAMS_HPP  // skip the 2.4 MB fixture header (memdata_nonsquare.h:1-2): PARAM:: is declared empty, weights are generated here
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
#include "conv_nonsquare_top.cpp"  // the reference top, unmodified (conv2d<>, and through bnn-library.h: dma.h, streamtools.h)

#define FCB_HLS_ADAPTER_QDMA
#include "finnconv_hls_adapter.hpp"

#include <cstdio>
#include <vector>

namespace {
// the case: K5 S2 P2, 16 -> 32 channels, 40 x 24 -> 20 x 12, SIMD 8, PE 8, 4-bit weights, 8-bit wrap + bias + ReLU
constexpr unsigned K = 5, S = 2, P = 2, C = 16, OFM = 32, IX = 40, IY = 24, OX = 20, OY = 12, SIMD = 8, PE = 8, WB = 4;
constexpr unsigned TILES = (K * K * C / SIMD) * (OFM / PE);
constexpr int WI = C * 8, WO = OFM * 8, DW = 64;
constexpr unsigned REPS = 19;  // 16 + 3: Mem2Stream_Batch / Stream2Mem_Batch take one 16-image burst and three single images

uint64_t rng_state = 0x1234abcdULL;
uint32_t rnd() {
  rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (uint32_t)(rng_state >> 33);
}

FixedPointWeights<SIMD, ap_int<WB>, PE, TILES> w;
FixedPointWeights<1, ap_int<8>, 1, OFM> b;

int checks = 0, failed = 0;
void expect(bool ok, const char* what) {
  checks++;
  if (!ok) { failed++; std::printf("adapter_check: FAILED %s\n", what); }
}
}  // namespace

int main() {
  for (unsigned pe = 0; pe < PE; pe++)
    for (unsigned t = 0; t < TILES; t++) w.m_weights[pe][t] = ap_uint<SIMD * WB>(rnd());
  for (unsigned o = 0; o < OFM; o++) b.m_weights[0][o] = ap_uint<8>(rnd() & 0xFF);
  std::vector<ap_uint<WI> > x((size_t)IX * IY * REPS);
  for (auto& v : x) {
    v = 0;
    for (int l = 0; l < (int)C; l++) v(8 * l + 7, 8 * l) = rnd() & 0x7F;
  }
  // ---- the reference, image by image (its numReps > 1 is not functional: SURVEY.md F7)
  std::vector<ap_uint<WO> > want;
  for (unsigned r = 0; r < REPS; r++) {
    hls::stream<ap_uint<WI> > si("si");
    hls::stream<ap_uint<WO> > so("so");
    for (size_t i = 0; i < (size_t)IX * IY; i++) si.write(x[r * (size_t)IX * IY + i]);
    conv2d<K, K, SIMD, PE, WB, C, OFM, IX, IY, OX, OY, S, S, P, 8, TILES, 8>(w, b, si, so, 1);
    while (!so.empty()) want.push_back(so.read());
  }
  expect(want.size() == (size_t)OX * OY * REPS, "reference output size");

  const fcb_layer_desc d = fcb_hls::layer_desc(FCB_KIND_CONV, K, S, P, C, OFM, IX, IY, SIMD, PE, WB);
  fcb_layer* L = fcb_hls::make_layer(d, w, b);

  {  // 1. HLS streams over one layer handle
    hls::stream<ap_uint<WI> > si("si");
    hls::stream<ap_uint<WO> > so("so");
    for (auto& v : x) si.write(v);
    fcb_hls::run_streams<WI, WO>(L, fcb_layer_run, si, so, REPS, (size_t)IX * IY, (size_t)OX * OY);
    bool ok = so.size() == want.size();
    for (size_t i = 0; ok && i < want.size(); i++) ok = so.read() == want[i];
    expect(ok, "run_streams(fcb_layer_run) == conv2d<>");
  }
  {  // 2. the same through a pool: every GPU of the box, and two replicas per GPU so that a one-GPU box still splits the batch
    std::vector<fcb_layer_desc> ds(1, d);
    std::vector<std::vector<uint8_t> > ws(1, fcb_hls::weight_image(w)), bs(1, fcb_hls::weight_image(b));
    fcb_pool* Pl = fcb_hls::make_pool(ds, ws, bs);
    const int ndev = (int)fcb_pool_replicas(Pl);
    fcb_pool_destroy(Pl);
    std::vector<int> devs;
    for (int i = 0; i < 2 * ndev; i++) devs.push_back(i % ndev);
    const void* wp[1] = {ws[0].data()};
    const void* bp[1] = {bs[0].data()};
    fcb_hls::check(fcb_pool_create(&d, wp, nullptr, bp, 1, devs.data(), (uint32_t)devs.size(), &Pl), "fcb_pool_create");
    hls::stream<ap_uint<WI> > si("si");
    hls::stream<ap_uint<WO> > so("so");
    for (auto& v : x) si.write(v);
    fcb_hls::run_streams<WI, WO>(Pl, fcb_pool_run, si, so, REPS, (size_t)IX * IY, (size_t)OX * OY);
    bool ok = so.size() == want.size();
    for (size_t i = 0; ok && i < want.size(); i++) ok = so.read() == want[i];
    expect(ok, "run_streams(fcb_pool_run) == conv2d<>");
    fcb_pool_destroy(Pl);
  }
  {  // 3. QDMA streams: the reference's adapters around its own layer give the expected output stream
    hls::stream<qdma_axis<WO, 0, 0, 0> > ref_q("ref_q");
    {
      hls::stream<ap_uint<WO> > so("so");
      for (auto& v : want) so.write(v);
      Stream2Qdma_Batch<WO, OX * OY>(so, ref_q, REPS);
    }
    hls::stream<qdma_axis<WI, 0, 0, 0> > qi("qi");
    hls::stream<qdma_axis<WO, 0, 0, 0> > qo("qo");
    {
      hls::stream<ap_uint<WI> > si("si");
      for (auto& v : x) si.write(v);
      Stream2Qdma_Batch<WI, IX * IY>(si, qi, REPS);  // a well-formed QDMA input stream, made by the reference
    }
    fcb_hls::run_qdma_streams<WI, WO>(L, fcb_layer_run, qi, qo, REPS, (size_t)IX * IY, (size_t)OX * OY);
    bool ok = qo.size() == ref_q.size();
    while (ok && !qo.empty()) {
      const qdma_axis<WO, 0, 0, 0> a = qo.read(), r = ref_q.read();
      ok = a.get_data() == r.get_data() && a.get_keep() == r.get_keep() && a.get_last() == r.get_last();
    }
    expect(ok, "run_qdma_streams == Qdma2Stream_Batch -> conv2d<> -> Stream2Qdma_Batch (data, TKEEP, TLAST)");
  }
  {  // 4. AXI memory: the reference's DMA blocks and width converters around its layer define the memory images
    constexpr unsigned IN_BYTES = IX * IY * WI / 8, OUT_BYTES = OX * OY * WO / 8;  // per image
    std::vector<ap_uint<DW> > in_mem((size_t)REPS * IN_BYTES / (DW / 8)), out_ref((size_t)REPS * OUT_BYTES / (DW / 8)), out_got(out_ref.size());
    {
      hls::stream<ap_uint<WI> > si("si");
      hls::stream<ap_uint<DW> > s64("s64");
      for (auto& v : x) si.write(v);
      StreamingDataWidthConverter_Batch<WI, DW, IX * IY>(si, s64, REPS);
      Stream2Mem_Batch<DW, IN_BYTES>(s64, in_mem.data(), REPS);  // the memory image a host would have prepared
    }
    {
      hls::stream<ap_uint<WO> > so("so");
      hls::stream<ap_uint<DW> > s64("s64");
      for (auto& v : want) so.write(v);
      StreamingDataWidthConverter_Batch<WO, DW, OX * OY>(so, s64, REPS);
      Stream2Mem_Batch<DW, OUT_BYTES>(s64, out_ref.data(), REPS);
    }
    {  // and Mem2Stream_Batch reads that image back as the stream the layer expects
      hls::stream<ap_uint<DW> > s64("s64");
      hls::stream<ap_uint<WI> > si("si");
      Mem2Stream_Batch<DW, IN_BYTES>(in_mem.data(), s64, REPS);
      StreamingDataWidthConverter_Batch<DW, WI, IX * IY * WI / DW>(s64, si, REPS);
      bool ok = si.size() == x.size();
      for (size_t i = 0; ok && i < x.size(); i++) ok = si.read() == x[i];
      expect(ok, "Mem2Stream_Batch -> DWC reproduces the input stream");
    }
    fcb_hls::run_axi_memory<WI, WO, DW>(L, fcb_layer_run, in_mem.data(), out_got.data(), REPS, (size_t)IX * IY, (size_t)OX * OY);
    bool ok = true;
    for (size_t i = 0; ok && i < out_ref.size(); i++) ok = out_got[i] == out_ref[i];
    expect(ok, "run_axi_memory == Mem2Stream_Batch -> DWC -> conv2d<> -> DWC -> Stream2Mem_Batch");
  }
  fcb_layer_destroy(L);
  if (failed) { std::printf("adapter_check: %d of %d checks FAILED\n", failed, checks); return 1; }
  std::printf("adapter_check: all %d checks passed\n", checks);
  return 0;
}
