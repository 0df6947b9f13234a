/*
 * ggml_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY; see ggml_oracle.c).
 * Type ids follow ggml_type, /root/reference/hf-kernels/ggml-kernels/ggml/ggml-common.h:1128-1161.
 */
#ifndef GGML_ORACLE_H
#define GGML_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_Q4_0 = 2, ORC_Q4_1 = 3, ORC_Q5_0 = 6, ORC_Q5_1 = 7, ORC_Q8_0 = 8, ORC_Q8_1 = 9,
    ORC_Q2_K = 10, ORC_Q3_K = 11, ORC_Q4_K = 12, ORC_Q5_K = 13, ORC_Q6_K = 14,
    ORC_IQ4_NL = 20, ORC_IQ4_XS = 23,
    ORC_IQ2_XXS = 16, ORC_IQ2_XS = 17, ORC_IQ3_XXS = 18, ORC_IQ1_S = 19, ORC_IQ3_S = 21, ORC_IQ2_S = 22, ORC_IQ1_M = 29
};

int     orc_block_elems(int type);
int     orc_block_bytes(int type);
int     orc_mmq_need_sum(int type);

/* fp32 dequantize of nelem values; single-threaded like the reference CPU op. */
int     orc_dequantize(const void *w, int type, int64_t nelem, float *out);
/* same, rows split over OpenMP threads (bench "all cores" leg; not the reference's behaviour). */
int     orc_dequantize_mt(const void *w, int type, int64_t rows, int64_t cols, float *out);

int64_t orc_q8_1_row_bytes(int64_t K);
int     orc_quantize_q8_1(const float *x, int64_t rows, int64_t K, void *out);
int64_t orc_q8_1_mmq_bytes(int64_t M, int64_t K);
int     orc_quantize_q8_1_mmq(const float *x, int64_t M, int64_t K, int need_sum, void *out);

/* fp16 bit patterns of the A operand the MMQ kernel builds (product-defined; see ggml_oracle.c) */
int     orc_mmq_a_operand(const void *w, int type, int64_t nelem, uint16_t *out);

int     orc_mmvq(const void *W, int type, const void *q8_1, int64_t B, int64_t K, int64_t N, float *Y);
int     orc_mmq(const void *W, int type, const void *q8_1_mmq, int64_t M, int64_t K, int64_t N, float *Y);

int     orc_num_threads(void);
int     orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
