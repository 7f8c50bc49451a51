// finnconv_hls_adapter.hpp -- header-only C++ adapter that puts the reference's HLS-stream interface on top
// of the C ABI (finnconv_b200.h), so code written against the reference top functions
//     void conv2d_layer0   (stream<ap_uint<3*8>>&,   stream<ap_uint<128*8>>&, unsigned)   conv_nonsquare_top.cpp:282
//     void deconv2d_layer4 (stream<ap_uint<192*8>>&, stream<ap_uint<128*8>>&, unsigned)   conv_nonsquare_top.cpp:288
//     void eight_layers_net(stream<ap_uint<24>>&,    stream<ap_uint<24>>&,    unsigned)   conv_nonsquare_top.cpp:295
// (e.g. the unmodified testbench conv3_nonsquare_tb.cpp) can link against the B200 backend instead of
// conv_nonsquare_top.cpp.  It needs ap_int.h / hls_stream.h / weights.hpp from the user's environment (the Xilinx
// headers, or oracle/shim in this repo); it contains no arithmetic of the layer: it drains the input stream into the
// packed word image the ABI takes, calls the GPU library and refills the output stream.
#ifndef FINNCONV_HLS_ADAPTER_HPP
#define FINNCONV_HLS_ADAPTER_HPP

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "finnconv_b200.h"

namespace fcb_hls {

inline void check(int rc, const char* what) {
  if (rc != FCB_OK) {
    // the reference's CASSERT_DATAFLOW prints and exits (bnn-library.h:55); keep that behaviour at this level only
    std::fprintf(stderr, "finnconv_b200: %s failed (%d): %s\n", what, rc, fcb_last_error());
    std::exit(-1);
  }
}

// ap_uint<W> <-> ap-word container (little endian, 1/2/4/8 bytes then multiples of 8)
template <int W> inline void word_to_bytes(const ap_uint<W>& v, uint8_t* p) {
  std::memset(p, 0, fcb_word_bytes(W));
  for (int b = 0; b < W; b += 8) {
    const int n = (W - b) < 8 ? (W - b) : 8;
    p[b / 8] = (uint8_t)(unsigned long long)v(b + n - 1, b);
  }
}
template <int W> inline ap_uint<W> bytes_to_word(const uint8_t* p) {
  ap_uint<W> v = 0;
  for (int b = 0; b < W; b += 8) {
    const int n = (W - b) < 8 ? (W - b) : 8;
    v(b + n - 1, b) = (unsigned long long)(p[b / 8] & ((1u << n) - 1u));
  }
  return v;
}

// memory image of FixedPointWeights::m_weights[PE][TILES] (weights.hpp:113)
template <unsigned SIMD, typename WT, unsigned PE, unsigned TILES>
inline std::vector<uint8_t> weight_image(const FixedPointWeights<SIMD, WT, PE, TILES>& w) {
  const size_t cb = fcb_word_bytes(SIMD * WT::width);
  std::vector<uint8_t> img(cb * PE * TILES);
  for (unsigned pe = 0; pe < PE; pe++)
    for (unsigned t = 0; t < TILES; t++) word_to_bytes<SIMD * WT::width>(w.m_weights[pe][t], &img[(size_t)(pe * TILES + t) * cb]);
  return img;
}

// descriptor of a conv2d<> / deconv522<> instantiation with the reference network's numerics
// (u8 activations, signed WIDTH-bit weights, PassThroughActivation<ap_uint<8>> + bias + ReLU)
inline fcb_layer_desc layer_desc(int kind, unsigned K, unsigned S, unsigned P, unsigned C, unsigned OFM, unsigned IX, unsigned IY,
                                 unsigned SIMD, unsigned PE, unsigned w_bits) {
  fcb_layer_desc d;
  std::memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.kind = kind;
  d.kernel_x = d.kernel_y = K;
  d.ifm_ch = C; d.ofm_ch = OFM; d.ifm_x = IX; d.ifm_y = IY;
  d.stride_x = d.stride_y = S; d.pad = P; d.simd = SIMD; d.pe = PE;
  d.ofm_x = kind == FCB_KIND_DECONV522 ? 2 * IX : (IX + 2 * P - K) / S + 1;
  d.ofm_y = kind == FCB_KIND_DECONV522 ? 2 * IY : (IY + 2 * P - K) / S + 1;
  d.in_bits = 8; d.in_signed = 0; d.w_bits = w_bits; d.weight_kind = FCB_W_FIXED;
  d.acc_bits = 8; d.acc_signed = 0; d.act_kind = FCB_ACT_BIAS_RELU; d.out_bits = 8;
  return d;
}

template <unsigned SIMD, typename WT, unsigned PE, unsigned TILES, unsigned OFM>
inline fcb_layer* make_layer(const fcb_layer_desc& d, const FixedPointWeights<SIMD, WT, PE, TILES>& w,
                             const FixedPointWeights<1, ap_int<8>, 1, OFM>& bias, int device = 0) {
  std::vector<uint8_t> wi = weight_image(w), bi = weight_image(bias);
  fcb_layer* L = nullptr;
  check(fcb_layer_create(&d, wi.data(), nullptr, bi.data(), device, &L), "fcb_layer_create");
  return L;
}

// page-locked staging buffer (fcb_host_alloc): the host-buffer calls copy from / to it at full PCIe rate
struct HostBuf {
  uint8_t* p = nullptr;
  explicit HostBuf(size_t bytes) {
    void* v = nullptr;
    check(fcb_host_alloc(&v, bytes), "fcb_host_alloc");
    p = (uint8_t*)v;
  }
  ~HostBuf() { fcb_host_free(p); }
  HostBuf(const HostBuf&) = delete;
  HostBuf& operator=(const HostBuf&) = delete;
};

// top(in, out, numReps): drain -> run -> refill.  RUN is fcb_layer_run, fcb_net_run or fcb_pool_run (the whole box behind one call).
template <int WI, int WO, typename H, typename RUN>
inline void run_streams(H* handle, RUN run, hls::stream<ap_uint<WI> >& in, hls::stream<ap_uint<WO> >& out, unsigned numReps,
                        size_t in_words_per_rep, size_t out_words_per_rep) {
  const size_t ib = fcb_word_bytes(WI), ob = fcb_word_bytes(WO);
  HostBuf hin(ib * in_words_per_rep * numReps), hout(ob * out_words_per_rep * numReps);
  for (size_t i = 0; i < in_words_per_rep * numReps; i++) word_to_bytes<WI>(in.read(), hin.p + i * ib);
  check(run(handle, hin.p, hout.p, numReps), "run");
  for (size_t i = 0; i < out_words_per_rep * numReps; i++) out.write(bytes_to_word<WO>(hout.p + i * ob));
}

#ifdef FCB_HLS_ADAPTER_QDMA
// The same with the Vitis top-level stream type on both sides: what Qdma2Stream_Batch in front of and Stream2Qdma_Batch behind the
// layer do (streamtools.h:1001-1037): data passes through, TKEEP is all ones, TLAST marks the last word of every frame.
// (needs ap_axi_sdata.h; define FCB_HLS_ADAPTER_QDMA before including this header)
template <int WI, int WO, typename H, typename RUN>
inline void run_qdma_streams(H* handle, RUN run, hls::stream<qdma_axis<WI, 0, 0, 0> >& in, hls::stream<qdma_axis<WO, 0, 0, 0> >& out,
                             unsigned numReps, size_t in_words_per_rep, size_t out_words_per_rep) {
  const size_t ib = fcb_word_bytes(WI), ob = fcb_word_bytes(WO);
  HostBuf hin(ib * in_words_per_rep * numReps), hout(ob * out_words_per_rep * numReps);
  for (size_t i = 0; i < in_words_per_rep * numReps; i++) word_to_bytes<WI>(ap_uint<WI>(in.read().get_data()), hin.p + i * ib);
  check(run(handle, hin.p, hout.p, numReps), "run");
  for (unsigned rep = 0; rep < numReps; rep++)
    for (size_t w = 0; w < out_words_per_rep; w++) {
      qdma_axis<WO, 0, 0, 0> t;
      t.set_data(bytes_to_word<WO>(hout.p + (rep * out_words_per_rep + w) * ob));
      t.set_keep(-1);
      t.set_last(w == out_words_per_rep - 1);
      out.write(t);
    }
}
#endif

// AXI-memory form: what Mem2Stream_Batch -> StreamingDataWidthConverter_Batch in front of and DWC -> Stream2Mem_Batch behind the layer
// do (dma.h:135-199, streamtools.h:463-526).  `in_mem` / `out_mem` are arrays of ap_uint<DW> memory words holding numReps frames back
// to back; both converters move bits LSB-first, so a frame is the dense bit string of its stream words (WI or WO bits each) cut into
// DW-bit memory words.  DW must divide or be a multiple of WI and WO, as the converters require.
template <int WI, int WO, int DW, typename H, typename RUN>
inline void run_axi_memory(H* handle, RUN run, const ap_uint<DW>* in_mem, ap_uint<DW>* out_mem, unsigned numReps, size_t in_words_per_rep,
                           size_t out_words_per_rep) {
  static_assert((WI % DW == 0 || DW % WI == 0) && (WO % DW == 0 || DW % WO == 0), "StreamingDataWidthConverter_Batch needs integer ratios");
  const size_t ib = fcb_word_bytes(WI), ob = fcb_word_bytes(WO);
  const size_t nin = in_words_per_rep * numReps, nout = out_words_per_rep * numReps;
  HostBuf hin(ib * nin), hout(ob * nout);
  for (size_t i = 0; i < nin; i++) {  // stream word i = bits [i*WI, (i+1)*WI) of the memory bit string
    ap_uint<WI> v = 0;
    for (int b = 0; b < WI; b++) {
      const size_t bit = i * (size_t)WI + b;
      v[b] = in_mem[bit / DW][(int)(bit % DW)];
    }
    word_to_bytes<WI>(v, hin.p + i * ib);
  }
  check(run(handle, hin.p, hout.p, numReps), "run");
  const size_t out_mem_words = (nout * (size_t)WO + DW - 1) / DW;
  for (size_t m = 0; m < out_mem_words; m++) out_mem[m] = 0;
  for (size_t i = 0; i < nout; i++) {
    const ap_uint<WO> v = bytes_to_word<WO>(hout.p + i * ob);
    for (int b = 0; b < WO; b++) {
      const size_t bit = i * (size_t)WO + b;
      out_mem[bit / DW][(int)(bit % DW)] = v[b];
    }
  }
}

// The reference network on every GPU of the box behind the reference's own signature: eight_layers_net(in, out, numReps)
// (conv_nonsquare_top.cpp:295).  `descs` / `weights` / `biases`: one entry per layer (layer_desc(), weight_image()).
inline fcb_pool* make_pool(const std::vector<fcb_layer_desc>& descs, const std::vector<std::vector<uint8_t> >& weights,
                           const std::vector<std::vector<uint8_t> >& biases) {
  std::vector<const void*> w, b;
  for (size_t i = 0; i < descs.size(); i++) { w.push_back(weights[i].data()); b.push_back(biases[i].data()); }
  fcb_pool* P = nullptr;
  check(fcb_pool_create(descs.data(), w.data(), nullptr, b.data(), (uint32_t)descs.size(), nullptr, 0, &P), "fcb_pool_create");
  return P;
}

}  // namespace fcb_hls
#endif
