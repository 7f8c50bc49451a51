// fcb_umma_common.h -- shared between the host glue (fcb_plan.cu) and the kernels (fcb_umma2.cu; fcb_umma_v1.cu in experiment builds).
#pragma once
#include <cuda.h>

#include "fcb_internal.h"

namespace fcb {
int umma_encode_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
int umma_encode_map_ex(CUtensorMap* m, void* base, int elem_bytes, int swizzle, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box);

struct Umma2Plan;
int umma2_plan_create(const Geom& g, const int8_t* d_w, const EpiParams& epi, int num_sms, Umma2Plan** out);
// bias_word >= 0: d_w carries the bias in byte 4*bias_word of every row and the producer writes a constant 1 there
int umma2_plan_create_thin(const Geom& g, const int8_t* d_w /*[CB*128][128]*/, const EpiParams& epi, int num_sms, int bias_word, Umma2Plan** out);
int umma2_plan_create_dthin(const Geom& g, const int8_t* d_w /*[9*cch*16][128]*/, const EpiParams& epi, int num_sms, Umma2Plan** out);
int umma2_plan_create_dcol(const Geom& g, const int8_t* d_w /*[cch*128][128]*/, const EpiParams& epi, int num_sms, Umma2Plan** out);
void umma2_plan_destroy(Umma2Plan* U);
int umma2_run(Umma2Plan* U, const void* d_in, void* d_out, int n_images, cudaStream_t st);
const char* umma2_describe(const Umma2Plan* U, char* buf, size_t n);

struct UmmaV1;  // first-generation per-tap-TMA kernel: experiment builds only (fcb_umma_v1.cu)
#ifdef FCB_EXPERIMENT
int umma_v1_eligible(const Geom& g);
int umma_v1_create(const Geom& g, int8_t* d_w /*[OFMpad][K]*/, const EpiParams& epi, int num_sms, UmmaV1** out);
void umma_v1_destroy(UmmaV1* p);
int umma_v1_run(UmmaV1* p, const void* d_in, void* d_out, int n_images, cudaStream_t st);
#endif
}  // namespace fcb
