// oracle/tb_b200_top.cpp -- TEST INFRASTRUCTURE ONLY (it includes reference headers from /root/reference).
//
// Stands in for conv_nonsquare_top.cpp when linking the reference's UNMODIFIED testbench
// (conv3_nonsquare_tb.cpp) against the B200 backend: defines conv2d_layer0 / deconv2d_layer4 /
// eight_layers_net with the reference's exact signatures (conv_nonsquare_top.cpp:282,288,295), built from the
// reference's own fixture weights (memdata_nonsquare.h) and parameters (config_nonsquare.h), on top of
// include/finnconv_hls_adapter.hpp + libfinnconv_b200.so.  `make -C oracle tb_b200` -> oracle/_ref/tb_b200; the
// testbench then prints "Image # 0 passed the testing." if the GPU network equals the testbench's own golden chain.
#define AP_INT_MAX_W 16384
#include <hls_stream.h>
using namespace hls;
#include "ap_int.h"
#include "weights.hpp"
#include "memdata_nonsquare.h"
#include "config_nonsquare.h"
#include "finnconv_hls_adapter.hpp"

#define LAYER(n, kind)                                                                                                       \
  fcb_hls::make_layer(fcb_hls::layer_desc(kind, CONV_##n##_K, CONV_##n##_S, CONV_##n##_P, CONV_##n##_IFM_CH, CONV_##n##_OFM_CH, \
                                          CONV_##n##_IFM_ROW, CONV_##n##_IFM_COL, CONV_##n##_SIMD, CONV_##n##_PE, CONV_##n##_W_BIT), \
                      PARAM::weights_layer##n, PARAM::bias_layer##n)

void conv2d_layer0(stream<ap_uint<CONV_0_IFM_CH * CONV_0_IN_BIT> >& in, stream<ap_uint<CONV_0_OFM_CH * CONV_0_OUT_BIT> >& out,
                   unsigned int numReps) {
  static fcb_layer* L = LAYER(0, FCB_KIND_CONV);
  fcb_hls::run_streams<CONV_0_IFM_CH * CONV_0_IN_BIT, CONV_0_OFM_CH * CONV_0_OUT_BIT>(
      L, fcb_layer_run, in, out, numReps, (size_t)CONV_0_IFM_ROW * CONV_0_IFM_COL, (size_t)CONV_0_OFM_ROW * CONV_0_OFM_COL);
}

void deconv2d_layer4(stream<ap_uint<CONV_4_IFM_CH * CONV_4_IN_BIT> >& in, stream<ap_uint<CONV_4_OFM_CH * CONV_4_OUT_BIT> >& out,
                     unsigned int numReps) {
  static fcb_layer* L = LAYER(4, FCB_KIND_DECONV522);
  fcb_hls::run_streams<CONV_4_IFM_CH * CONV_4_IN_BIT, CONV_4_OFM_CH * CONV_4_OUT_BIT>(
      L, fcb_layer_run, in, out, numReps, (size_t)CONV_4_IFM_ROW * CONV_4_IFM_COL, (size_t)CONV_4_OFM_ROW * CONV_4_OFM_COL);
}

void eight_layers_net(stream<ap_uint<CONV_0_IFM_CH * CONV_0_IN_BIT> >& in, stream<ap_uint<CONV_7_OFM_CH * CONV_7_OUT_BIT> >& out,
                      unsigned int numReps) {
  // every sm_100 GPU of the box behind the one call: numReps is split into contiguous image ranges (fcb_pool_*)
  static fcb_pool* N = []() {
    std::vector<fcb_layer_desc> d;
    std::vector<std::vector<uint8_t> > w, b;
#define ADD(n, kind)                                                                                                                   \
  d.push_back(fcb_hls::layer_desc(kind, CONV_##n##_K, CONV_##n##_S, CONV_##n##_P, CONV_##n##_IFM_CH, CONV_##n##_OFM_CH, CONV_##n##_IFM_ROW,  \
                                  CONV_##n##_IFM_COL, CONV_##n##_SIMD, CONV_##n##_PE, CONV_##n##_W_BIT));                                 \
  w.push_back(fcb_hls::weight_image(PARAM::weights_layer##n));                                                                          \
  b.push_back(fcb_hls::weight_image(PARAM::bias_layer##n));
    ADD(0, FCB_KIND_CONV) ADD(1, FCB_KIND_CONV) ADD(2, FCB_KIND_CONV) ADD(3, FCB_KIND_CONV)
    ADD(4, FCB_KIND_DECONV522) ADD(5, FCB_KIND_DECONV522) ADD(6, FCB_KIND_DECONV522) ADD(7, FCB_KIND_DECONV522)
#undef ADD
    return fcb_hls::make_pool(d, w, b);
  }();
  fcb_hls::run_streams<CONV_0_IFM_CH * CONV_0_IN_BIT, CONV_7_OFM_CH * CONV_7_OUT_BIT>(
      N, fcb_pool_run, in, out, numReps, (size_t)CONV_0_IFM_ROW * CONV_0_IFM_COL, (size_t)CONV_7_OFM_ROW * CONV_7_OFM_COL);
}
