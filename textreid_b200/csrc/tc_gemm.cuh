// bf16 tcgen05 GEMM building blocks of the loss path (tc_gemm.cu).
#pragma once
#include "common.cuh"

// bytes of the packed bf16 image of a [rows, kdim] operand (rows padded to 256, kdim to 64)
int64_t tc_gemm_packed_bytes(int rows, int kdim);
// element (r, k) of the operand is src[r * s_row + k * s_k]; optional per-k / per-row scale before the rounding to bf16
int tc_pack_strided(const float* src, int64_t s_row, int64_t s_k, int rows, int kdim, const float* kscale, const float* rscale,
                    void* packed, cudaStream_t st);
// C[ks][m, n] = nscale[n] * sum_{k in split ks} A[m, k] * B[n, k];  C row stride ldc, split stride c_split (elements)
int tc_gemm_launch(const void* A_packed, const void* B_packed, float* C, int64_t ldc, int64_t c_split, int M, int N, int kdim,
                   int ksplit, const float* nscale, cudaStream_t st);
