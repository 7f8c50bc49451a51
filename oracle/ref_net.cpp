// oracle/ref_net.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference top conv_nonsquare_top.cpp, UNMODIFIED and with its own fixture weights
// (memdata_nonsquare.h), compiled where it lies under /root/reference into
// oracle/_ref/libref_net.so, plus C wrappers so that Python can (a) run eight_layers_net /
// conv2d_layer0 / deconv2d_layer4 (conv_nonsquare_top.cpp:282,288,295) on arbitrary images and
// (b) dump the PARAM:: weight and bias images for tests/golden/ (generated data, not source).
#include <cstdint>
#include <cstring>
#include <chrono>
#include "/root/reference/conv_nonsquare_top.cpp"

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
template <int WI, int WO, typename F>
int run_top(F fn, const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out, double* secs) {
  hls::stream<ap_uint<WI> > s_in("in");
  hls::stream<ap_uint<WO> > s_out("out");
  for (size_t i = 0; i < n_in; i++) s_in.write(load_word<WI>(in + i * container_bytes(WI)));
  auto t0 = std::chrono::steady_clock::now();
  fn(s_in, s_out, 1u);
  auto t1 = std::chrono::steady_clock::now();
  if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
  if (s_out.size() != n_out) return -2;
  for (size_t i = 0; i < n_out; i++) store_word<WO>(out + i * container_bytes(WO), s_out.read());
  return 0;
}
template <unsigned SIMD, typename WT, unsigned PE, unsigned TILES>
size_t dump(const FixedPointWeights<SIMD, WT, PE, TILES>& w, uint8_t* out) {
  const int cb = container_bytes(SIMD * WT::width);
  if (out)
    for (unsigned pe = 0; pe < PE; pe++)
      for (unsigned t = 0; t < TILES; t++) store_word<SIMD * WT::width>(out + (size_t)(pe * TILES + t) * cb, w.m_weights[pe][t]);
  return (size_t)PE * TILES * cb;
}
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int ref_eight_layers_net(const uint8_t* in, uint8_t* out, double* secs) {
  return run_top<CONV_0_IFM_CH * CONV_0_IN_BIT, CONV_7_OFM_CH * CONV_7_OUT_BIT>(
      eight_layers_net, in, (size_t)CONV_0_IFM_ROW * CONV_0_IFM_COL, out, (size_t)CONV_7_OFM_ROW * CONV_7_OFM_COL, secs);
}
REF_API int ref_conv2d_layer0(const uint8_t* in, uint8_t* out, double* secs) {
  return run_top<CONV_0_IFM_CH * CONV_0_IN_BIT, CONV_0_OFM_CH * CONV_0_OUT_BIT>(
      conv2d_layer0, in, (size_t)CONV_0_IFM_ROW * CONV_0_IFM_COL, out, (size_t)CONV_0_OFM_ROW * CONV_0_OFM_COL, secs);
}
REF_API int ref_deconv2d_layer4(const uint8_t* in, uint8_t* out, double* secs) {
  return run_top<CONV_4_IFM_CH * CONV_4_IN_BIT, CONV_4_OFM_CH * CONV_4_OUT_BIT>(
      deconv2d_layer4, in, (size_t)CONV_4_IFM_ROW * CONV_4_IFM_COL, out, (size_t)CONV_4_OFM_ROW * CONV_4_OFM_COL, secs);
}
// packed images of PARAM::weights_layerN / bias_layerN; returns byte count (out may be NULL)
REF_API long ref_dump_weights(int layer, uint8_t* out) {
  switch (layer) {
    case 0: return (long)dump(PARAM::weights_layer0, out);
    case 1: return (long)dump(PARAM::weights_layer1, out);
    case 2: return (long)dump(PARAM::weights_layer2, out);
    case 3: return (long)dump(PARAM::weights_layer3, out);
    case 4: return (long)dump(PARAM::weights_layer4, out);
    case 5: return (long)dump(PARAM::weights_layer5, out);
    case 6: return (long)dump(PARAM::weights_layer6, out);
    case 7: return (long)dump(PARAM::weights_layer7, out);
  }
  return -1;
}
REF_API long ref_dump_bias(int layer, uint8_t* out) {
  switch (layer) {
    case 0: return (long)dump(PARAM::bias_layer0, out);
    case 1: return (long)dump(PARAM::bias_layer1, out);
    case 2: return (long)dump(PARAM::bias_layer2, out);
    case 3: return (long)dump(PARAM::bias_layer3, out);
    case 4: return (long)dump(PARAM::bias_layer4, out);
    case 5: return (long)dump(PARAM::bias_layer5, out);
    case 6: return (long)dump(PARAM::bias_layer6, out);
    case 7: return (long)dump(PARAM::bias_layer7, out);
  }
  return -1;
}
