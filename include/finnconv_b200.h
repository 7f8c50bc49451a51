/* finnconv_b200.h -- C ABI of the B200-native FINN-style quantized non-square
 * convolution layer (sliding window + MVAU MAC array + activation [+ pool]).
 *
 * This header is the drop-in boundary for the reference's HLS top functions.
 * The reference has no FFI layer; its boundary is the HLS top-function
 * signature over hls::stream<ap_uint<W>> (conv_nonsquare_top.cpp:282,288,295)
 * and the template parameter sets of conv2d<> (conv_nonsquare_top.cpp:198-215),
 * deconv522<> (:71-81), ConvLayer_Batch (convlayer.h:89-111) and
 * ThresholdsActivation (activations.hpp:168-169).  Every entry point below
 * names the reference interface it replaces.
 *
 * DATA LAYOUTS (accepted verbatim):
 *  - "ap-word container": the memory image of one ap_uint<W>/ap_int<W>:
 *    little-endian, 1 byte (W<=8), 2 (W<=16), 4 (W<=32), 8 (W<=64), else
 *    8*ceil(W/64) bytes.  Only the low W bits are read; writers zero the rest.
 *  - stream: numReps images back to back; per image ifm_y rows (y-major) of
 *    ifm_x pixels (x fastest); one word ap_uint<ifm_ch*in_bits> per pixel;
 *    channel c occupies bits [c*in_bits, (c+1)*in_bits) (lane 0 at the LSB:
 *    interpret.hpp:211, conv3_nonsquare_tb.cpp:807-808).
 *  - weights: image of FixedPointWeights::m_weights[PE][TILES]
 *    (weights.hpp:113) -- words ap_uint<SIMD*w_bits>, pe-major; lane `simd` of
 *    [pe][tile] is W[ch = nf*PE + pe][k = sf*SIMD + simd], tile = nf*SF + sf,
 *    k = (ky*Kx + kx)*ifm_ch + c (mvau.hpp:101-148, slidingwindow.h:1302-1313).
 *    BinaryWeights::m_weights[PE][TILES] (weights.hpp:69): words ap_uint<SIMD>.
 *  - thresholds: image of ThresholdsActivation::m_thresholds[PE][NF][NumTH]
 *    (activations.hpp:172), each an ap-word container of acc_bits.
 *  - bias: image of FixedPointWeights<1,ap_int<8>,1,OFM>::m_weights[1][OFM]
 *    (memdata_nonsquare.h:16-22): ofm_ch bytes, signed.
 *
 * All functions return FCB_OK (0) or a negative fcb_status; the message of the
 * last failure on the calling thread is available through fcb_last_error().
 * The reference instead prints and exit(-1)s (CASSERT_DATAFLOW,
 * bnn-library.h:55); this library never exits the process.
 *
 * There is NO CPU fallback: every run entry point executes CUDA kernels built
 * for sm_100a and fails with FCB_ERR_CUDA when no such device is usable.
 */
#ifndef FINNCONV_B200_H
#define FINNCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FCB_API
#else
#define FCB_API __attribute__((visibility("default")))
#endif

typedef enum fcb_status {
  FCB_OK = 0,
  FCB_ERR_INVALID_ARG = -1, /* null pointer, bad struct_size, bad enum */
  FCB_ERR_SHAPE = -2,       /* violates a reference CASSERT (IFM_CH % SIMD, OFM_CH % PE, geometry) */
  FCB_ERR_UNSUPPORTED = -3, /* valid in the reference, not implemented here */
  FCB_ERR_CUDA = -4,        /* CUDA runtime/driver failure, or no sm_100 device */
  FCB_ERR_NOMEM = -5
} fcb_status;

/* conv2d<> (conv_nonsquare_top.cpp:198-280) vs deconv522<> (:71-195) */
typedef enum fcb_layer_kind {
  FCB_KIND_CONV = 0,
  FCB_KIND_DECONV522 = 1,
  /* depth-wise convolution: [FMPadding_nonsquare ->] ConvolutionInputGenerator[_NonSquare]_dws (slidingwindow.h:761-868, 1377-1488)
   * -> Vector_Vector_Activate_Batch (vvau.hpp:80-154).  ofm_ch == ifm_ch, simd == pe (the generator's SIMD is the VVAU's PE);
   * weights = image of FixedPointWeights<1, ap_int<w_bits>, PE, NF*Kx*Ky>::m_weights[PE][TILES], one lane per ap_uint<w_bits> word,
   * tile = nf*Kx*Ky + ky*Kx + kx (the generator emits, per channel chunk, the taps in (ky, kx) order); any activation. */
  FCB_KIND_DWCONV = 2,
  /* generic pooling: the same sliding window -> Pool_batch (maxpool.h:525-577) with a pool.hpp function (pool.hpp:94-226).
   * ofm_ch == ifm_ch, simd == pe; no weights (NULL); `weight_kind` holds the fcb_pool_fn, `act_val` its `size` argument (divisor of
   * AvgPoolFunction, shift of QuantAvgPoolFunction); in_bits / in_signed = TSrcI lanes, acc_bits / acc_signed = the function's
   * accumulator type, out_bits = TDstI lanes; act_kind must be FCB_ACT_PASSTHROUGH; kernel / stride / padding as for FCB_KIND_CONV. */
  FCB_KIND_POOL = 3
} fcb_layer_kind;
/* MaxPoolFunction (init = type minimum) | AvgPoolFunction (sum, then / size, C++ truncation) | AccPoolFunction (sum) |
 * QuantAvgPoolFunction (sum, then >> size) -- pool.hpp:94-226 */
typedef enum fcb_pool_fn { FCB_POOLFN_MAX = 0, FCB_POOLFN_AVG = 1, FCB_POOLFN_ACC = 2, FCB_POOLFN_QUANTAVG = 3 } fcb_pool_fn;

/* weights.hpp:110-150 FixedPointWeights | weights.hpp:66-98 BinaryWeights with
 * Recast<XnorMul> activations (interpret.hpp:57-73) | BinaryWeights with
 * Recast<Binary> weights, +-1 (interpret.hpp:75-108) */
typedef enum fcb_weight_kind { FCB_W_FIXED = 0, FCB_W_BINARY_XNOR = 1, FCB_W_BINARY_PM1 = 2 } fcb_weight_kind;

/* activations.hpp:127-134 | conv_nonsquare_top.cpp:267-278 | activations.hpp:168-190 */
typedef enum fcb_act_kind { FCB_ACT_PASSTHROUGH = 0, FCB_ACT_BIAS_RELU = 1, FCB_ACT_THRESHOLDS = 2 } fcb_act_kind;

/* comp::less (default: thr < acc), greater, less_equal, greater_equal -- activations.hpp:57-99;
 * evaluated as Compare()(threshold, accu) (activations.hpp:185) */
typedef enum fcb_cmp { FCB_CMP_LESS = 0, FCB_CMP_GREATER = 1, FCB_CMP_LESS_EQUAL = 2, FCB_CMP_GREATER_EQUAL = 3 } fcb_cmp;

/* Which arithmetic unit multiplies: the counterpart of the reference's resource argument `R` of
 * Matrix_Vector_Activate_Batch (ap_resource_dsp / ap_resource_lut / ap_resource_dflt, mvau.hpp:87-98, mac.hpp:87-144), which
 * like here changes the implementation and never the result.  AUTO picks the fastest engine that covers the layer;
 * IMAD = CUDA-core integer MACs (covers everything); XNOR_POPC = XNOR + __popc on packed 1-bit lanes (FCB_W_BINARY_XNOR only);
 * TENSOR = tcgen05 kind::i8 (for FCB_W_BINARY_XNOR: lanes expanded to +-1 int8, thresholds remapped 2t-K).  A hint the layer
 * cannot honour fails fcb_layer_create with FCB_ERR_UNSUPPORTED. */
typedef enum fcb_engine_hint { FCB_ENGINE_AUTO = 0, FCB_ENGINE_IMAD = 1, FCB_ENGINE_XNOR_POPC = 2, FCB_ENGINE_TENSOR = 3 } fcb_engine_hint;

/* Run-time mirror of the reference's compile-time parameter set. */
typedef struct fcb_layer_desc {
  uint32_t struct_size; /* = sizeof(fcb_layer_desc) */
  uint32_t kind;        /* fcb_layer_kind */
  /* geometry: KERNEL_DIM_X/Y, IFM/OFM_Channels, IFMDim_x/y, OFMDim_x/y, STRIDE_x/y, PADDING */
  uint32_t kernel_x, kernel_y;
  uint32_t ifm_ch, ofm_ch;
  uint32_t ifm_x, ifm_y; /* x = fast (row-of-pixels) axis, the reference's "ROW" (768-side) */
  uint32_t ofm_x, ofm_y; /* conv output extent BEFORE pooling; must equal the geometry's */
  uint32_t stride_x, stride_y;
  uint32_t pad; /* zeros added on each side (FMPadding total = 2*pad, streamtools.h:374-379) */
  /* folding (only defines the weight/threshold image layout) */
  uint32_t simd, pe;
  /* numerics */
  uint32_t in_bits;     /* INPUT_PRECISION */
  uint32_t in_signed;   /* TSrcI = Slice<ap_int<>> (1) or Slice<ap_uint<>> (0) */
  uint32_t w_bits;      /* WIDTH of FixedPointWeights lanes; 1 for binary kinds */
  uint32_t weight_kind; /* fcb_weight_kind */
  uint32_t acc_bits;    /* TA = decltype(activation.init(0,0)) width (mvau.hpp:112) */
  uint32_t acc_signed;  /* TA signedness */
  uint32_t act_kind;    /* fcb_act_kind */
  uint32_t out_bits;    /* TDstI lane width / ACTIVATION_PRECISION / TR width (unsigned) */
  uint32_t num_th;      /* NumTH (thresholds only) */
  int32_t act_val;      /* ActVal */
  uint32_t cmp;         /* fcb_cmp */
  uint32_t pool;        /* 0 or 1: none; k>=2: k x k, stride k max pool on the out_bits lanes
                           (StreamingMaxPool_Precision, maxpool.h:137-185; OR for out_bits==1, :66-96) */
  uint32_t engine_hint; /* fcb_engine_hint; 0 = automatic */
  /* FMPadding_nonsquare's own parameter set (streamtools.h:361-379) for paddings `pad` cannot express: used when
   * pad_style != 0 (then `pad` must be 0).  Padding_x = pad_x_total zeros are split left = P/2 + (pad_style == 2 ? P % 2 : 0),
   * right = P - left; likewise up / down with pad_y_total.  pad_style 2 is the reference's default PaddingStyle. */
  uint32_t pad_x_total, pad_y_total, pad_style;
  /* StreamingMaxPool_Precision's ActType signedness and min_value (maxpool.h:137-170): lanes are compared as signed
   * out_bits-wide integers when pool_signed != 0; every window's maximum starts from pool_min_value. */
  uint32_t pool_signed;
  int32_t pool_min_value;
  /* Dilation_x / Dilation_y of ConvolutionInputGenerator_NonSquare_Dilated (slidingwindow.h:1515-1631): tap (ky, kx) reads the padded
   * frame at (oy*S + ky*Dy, ox*S + kx*Dx); 0 or 1 = none.  (The reference asserts Dilation_y == 1; both axes work here.)
   * These two fields were appended in ABI 0.2: a struct_size without them (FCB_LAYER_DESC_SIZE_V1) is accepted and means 1. */
  uint32_t dilation_x, dilation_y;
} fcb_layer_desc;
#define FCB_LAYER_DESC_SIZE_V1 (sizeof(fcb_layer_desc) - 2 * sizeof(uint32_t))

typedef struct fcb_layer fcb_layer; /* opaque: one layer resident on one device */
typedef struct fcb_net fcb_net;     /* opaque: chain of layers, activations stay on device */

/* --- library ---------------------------------------------------------------- */
FCB_API const char* fcb_version(void);
FCB_API const char* fcb_last_error(void);
/* number of usable sm_100 devices (0 if none); never fails */
FCB_API int fcb_device_count(void);
/* bytes of one ap-word container of `bits` bits (1,2,4,8,16,24,...) */
FCB_API size_t fcb_word_bytes(uint32_t bits);
/* Validate a descriptor exactly as the reference's CASSERTs would (slidingwindow.h:1259,
 * streamtools.h:472,505, mvau.hpp:101-105) and report the stream sizes. Any out pointer may be NULL. */
FCB_API int fcb_layer_query(const fcb_layer_desc* desc, size_t* in_bytes_per_image, size_t* out_bytes_per_image,
                            size_t* weight_bytes, size_t* threshold_bytes, size_t* bias_bytes);

/* --- one layer: replaces conv2d<>/deconv522<> instantiations such as conv2d_layer0
 * (conv_nonsquare_top.cpp:282) and deconv2d_layer4 (:288), and ConvLayer_Batch (convlayer.h:106-125).
 * `weights` is the m_weights image, `thresholds` the m_thresholds image (FCB_ACT_THRESHOLDS) and
 * `bias` the bias image (FCB_ACT_BIAS_RELU); unused ones may be NULL. Images are copied and
 * re-laid-out on `device`; the caller may free them on return. */
FCB_API int fcb_layer_create(const fcb_layer_desc* desc, const void* weights, const void* thresholds, const void* bias,
                             int device, fcb_layer** out);
FCB_API void fcb_layer_destroy(fcb_layer* layer);
/* Replace the layer's parameters in place (same descriptor, same handle): the run-time-writable weight memories of the
 * reference -- GenParamStream feeding Matrix_Vector_Activate_Stream_Batch (dma.h:214-236, mvau.hpp:209-307) -- as one call.
 * Images as for fcb_layer_create; on error the layer keeps its previous parameters. Not concurrent with runs of the layer. */
FCB_API int fcb_layer_set_params(fcb_layer* layer, const void* weights, const void* thresholds, const void* bias);
/* The same, with the weights in the reference's parameter-STREAM format: one period (TILES words) of what GenParamStream
 * (dma.h:214-236) writes and Matrix_Vector_Activate_Stream_Batch (mvau.hpp:209-307) reads -- word `tile` is an
 * ap_uint<SIMD*PE*WP> container holding m_weights[pe][tile] at bits [pe*SIMD*WP, (pe+1)*SIMD*WP). thresholds / bias as above. */
FCB_API int fcb_layer_set_param_stream(fcb_layer* layer, const void* param_words, const void* thresholds, const void* bias);
/* Host-buffer call: in_words -> H2D -> kernels -> D2H -> out_words, numReps images, synchronous.
 * Mirrors `top(in_stream, out_stream, numReps)`. */
FCB_API int fcb_layer_run(fcb_layer* layer, const void* in_words, void* out_words, uint32_t numReps);
/* Device-buffer call on `stream` (a cudaStream_t, or NULL for the default stream); asynchronous.
 * d_in/d_out hold the same word images in device memory of layer's device. */
FCB_API int fcb_layer_run_device(fcb_layer* layer, const void* d_in, void* d_out, uint32_t numReps, void* stream);
/* Every entry point runs on the handle's device and restores the caller's current CUDA device before it returns. */
/* Images per staging slot of fcb_layer_run (two slots double-buffer H2D / kernels / D2H); 0 restores the default
 * (<= 256 MiB per slot).  The burst length of the reference's Mem2Stream_Batch / Stream2Mem_Batch (16 images, dma.h:166-176). */
FCB_API int fcb_layer_set_host_chunk(fcb_layer* layer, uint32_t images);
/* Which kernel family serves this layer: "umma_i8", "imad", "xnor_popc" (diagnostics / tests). */
FCB_API const char* fcb_layer_engine(const fcb_layer* layer);
/* Human-readable tiling plan of the layer (tile shape, shared-memory planes, pipeline depth). */
FCB_API const char* fcb_layer_plan(const fcb_layer* layer);
/* Number of kernel launches issued by this layer since creation. */
FCB_API uint64_t fcb_layer_launches(const fcb_layer* layer);

/* --- layer chain: replaces eight_layers_net (conv_nonsquare_top.cpp:295-357).
 * Layer i's output word width must equal layer i+1's input word width. */
FCB_API int fcb_net_create(fcb_layer* const* layers, uint32_t n_layers, fcb_net** out); /* borrows the layers */
FCB_API void fcb_net_destroy(fcb_net* net);
FCB_API int fcb_net_run(fcb_net* net, const void* in_words, void* out_words, uint32_t numReps);
FCB_API int fcb_net_run_device(fcb_net* net, const void* d_in, void* d_out, uint32_t numReps, void* stream);
FCB_API uint64_t fcb_net_launches(const fcb_net* net);
/* Images per staging slot of fcb_net_run (0 = default, <= 64 MiB per slot), and images per pass of the layer chain inside
 * fcb_net_run_device (0 = default: the largest intermediate stream stays <= 1 GiB; small values keep the intermediates in L2). */
FCB_API int fcb_net_set_host_chunk(fcb_net* net, uint32_t images);
FCB_API int fcb_net_set_device_chunk(fcb_net* net, uint32_t images);

/* --- the whole box behind one handle: the chain (n_layers >= 1; one layer = conv2d_layer0-style tops) replicated on `devices`
 * (NULL / 0 = every sm_100 device; a device may be listed more than once) and numReps split into contiguous image ranges, one per
 * replica, each served by its own host thread, streams and staging slots.  eight_layers_net(in, out, numReps)
 * (conv_nonsquare_top.cpp:295) with numReps = 8*B then uses all 8 GPUs from ONE call; images are independent (the sliding-window
 * buffers reset per image, slidingwindow.h:1320,1351), weights are replicated, no device talks to another.
 * weights / thresholds / biases: one image pointer per layer (thresholds / biases may be NULL, or hold NULLs for layers without). */
typedef struct fcb_pool fcb_pool;
FCB_API int fcb_pool_create(const fcb_layer_desc* descs, const void* const* weights, const void* const* thresholds,
                            const void* const* biases, uint32_t n_layers, const int* devices, uint32_t n_devices, fcb_pool** out);
FCB_API void fcb_pool_destroy(fcb_pool* pool);
FCB_API uint32_t fcb_pool_replicas(const fcb_pool* pool);
/* host buffers in, host buffers out, synchronous; replica r takes images fcb_shard_range(numReps, r, replicas) */
FCB_API int fcb_pool_run(fcb_pool* pool, const void* in_words, void* out_words, uint32_t numReps);
/* the split itself: contiguous ranges, sizes differ by at most one image, the remainder goes to the low ranks */
FCB_API int fcb_shard_range(uint32_t numReps, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* end);
/* page-locked host memory, visible to every device: the host-buffer calls copy from / to it at full PCIe rate */
FCB_API int fcb_host_alloc(void** ptr, size_t bytes);
FCB_API void fcb_host_free(void* ptr);

/* --- residual add: AddStreams_Batch / AddStreamsLayer_Batch (streamtools.h:669-762).  Per stream word and channel
 * Out_t sum = op1 + op2 + offset with op1 / op2 read as In1_t / In2_t (ap_int / ap_uint of in*_bits) and the sum wrapped to out_bits.
 * Word images as everywhere: ap_uint<channels*bits> containers, n_words = NumTotal * numReps. */
typedef struct fcb_add_desc {
  uint32_t struct_size; /* = sizeof(fcb_add_desc) */
  uint32_t channels;
  uint32_t in1_bits, in1_signed, in2_bits, in2_signed; /* 1..32 bits */
  uint32_t out_bits;                                   /* 1..32 */
  int32_t offset;
} fcb_add_desc;
/* device buffers on `device`, asynchronous on `stream` */
FCB_API int fcb_add_streams_device(const fcb_add_desc* desc, const void* d_in1, const void* d_in2, void* d_out, uint64_t n_words, int device,
                                   void* stream);
/* host buffers, synchronous */
FCB_API int fcb_add_streams(const fcb_add_desc* desc, const void* in1, const void* in2, void* out, uint64_t n_words, int device);

/* --- synthetic data (bench / tests): byte i of the buffer = splitmix64(seed ^ (offset + i)) & mask,
 * the rule of SURVEY.md 8(d); d_ptr is device memory, 16-byte aligned; asynchronous on `stream`. */
FCB_API int fcb_synth_fill(void* d_ptr, size_t n_bytes, uint64_t seed, uint32_t mask, uint64_t offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FINNCONV_B200_H */
