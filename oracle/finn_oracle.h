/* oracle/finn_oracle.h -- TEST INFRASTRUCTURE ONLY (see finn_oracle.c).
 * CPU restatement of the reference layer; takes the same descriptor and the same packed
 * byte images as include/finnconv_b200.h so it can be fed the CUDA library's exact inputs. */
#ifndef FINN_ORACLE_H
#define FINN_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/finnconv_b200.h" /* fcb_layer_desc + enums only (interface, no product code) */

#ifdef __cplusplus
extern "C" {
#endif
#define FO_API __attribute__((visibility("default")))

typedef struct fo_sizes {
  uint32_t k_total, sf, nf, out_x, out_y;
  size_t in_word_bytes, out_word_bytes, in_bytes_per_image, out_bytes_per_image;
  size_t weight_word_bytes, weight_bytes, threshold_bytes, bias_bytes;
} fo_sizes;

FO_API size_t fo_word_bytes(uint32_t bits);
FO_API void fo_set_threads(int n); /* OpenMP threads over output rows (1 = scalar port) */
FO_API int fo_get_threads(void);
FO_API int fo_layer_query(const fcb_layer_desc* d, fo_sizes* s);
FO_API int fo_layer_run(const fcb_layer_desc* d, const void* in_words, const void* weights, const void* thresholds,
                        const void* bias, void* out_words, uint32_t numReps);
FO_API int fo_maxpool(const void* in_words, void* out_words, uint32_t dim_x, uint32_t dim_y, uint32_t pool, uint32_t ch,
                      uint32_t bits);
FO_API int fo_add_streams(const void* in1, const void* in2, void* out, uint64_t n_words, uint32_t channels, uint32_t in1_bits, int in1_signed,
                          uint32_t in2_bits, int in2_signed, uint32_t out_bits, int32_t offset);
#ifdef __cplusplus
}
#endif
#endif
