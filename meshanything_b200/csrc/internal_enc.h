// internal_enc.h -- launch helpers of glue.cu (encoder / detokenizer glue kernels)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ma {
int launch_fourier_embed(const __half* pc, long rows, __half* out, cudaStream_t st);
int launch_scatter_heads(const __half* src, int ld, int col0, int head_stride, int H, int rows_per_slot, long T,
                         __half* dst, long rows, cudaStream_t st);
int launch_residual_add(float* x32, __half* x16, const __half* y, long n, cudaStream_t st);
int launch_convert_rows(const void* src, int src_f16, long lds, void* dst, int dst_f16, long ldd, long rows, int cols,
                        long src_rows_mod, cudaStream_t st);
int launch_add_table(const __half* y16, const int* mask, const float* table, int table_rows, float* out, long rows,
                     cudaStream_t st);
int launch_gather_codes(const int32_t* gen_ids, int max_new, int B, int F, const float* codebook, __half* code16,
                        int* mask, int32_t* ids_out, cudaStream_t st);
int launch_coords(const __half* logits, const int* mask, float* xyz, long faces, cudaStream_t st);
}  // namespace ma
