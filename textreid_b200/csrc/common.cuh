// Shared device/host helpers for libtextreid_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math_constants.h>

#include "../../include/textreid_b200.h"

#define TRB_TOPK 10

// ---------------------------------------------------------------------------
// error plumbing: every entry point returns 0 or a negative trb code / cudaError
// ---------------------------------------------------------------------------
void trb_set_error(const char* fmt, ...);

#define TRB_REQUIRE(cond, ...)                                                   \
    do {                                                                         \
        if (!(cond)) {                                                           \
            trb_set_error(__VA_ARGS__);                                          \
            return TRB_ERR_INVALID;                                              \
        }                                                                        \
    } while (0)

#define TRB_CUDA_OK(expr)                                                        \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) {                                                 \
            trb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                      \
        }                                                                        \
    } while (0)

#define TRB_LAUNCH_OK() TRB_CUDA_OK(cudaGetLastError())

// cudaFuncSetAttribute is a PER-DEVICE setting: a call site remembers it per device ordinal (a zero-initialised static of this
// type), not per process.  Setting an attribute twice from two host threads is harmless.
struct TrbDeviceOnce { unsigned char done[64]; };
static inline bool trb_first_on_device(TrbDeviceOnce& once) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (once.done[dev]) return false;
    once.done[dev] = 1;
    return true;
}

static inline bool trb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t trb_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; every thread gets the result.  `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -CUDART_INF_F;
    r = warp_max(r);
    return r;
}

// Ranking order pinned by the north star: similarity descending, gallery index ascending.
template <typename T>
__device__ __forceinline__ bool ranks_before(T s_a, int64_t i_a, T s_b, int64_t i_b) {
    return (s_a > s_b) || (s_a == s_b && i_a < i_b);
}

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return -CUDART_INF_F; }
template <> __device__ __forceinline__ double neg_inf<double>() { return -CUDART_INF; }

// Per-thread best-TRB_TOPK list kept sorted (best first) in registers.  All indexing is static
// after unrolling, so the arrays stay in registers.
template <typename T>
struct TopKT {
    T s[TRB_TOPK];
    int64_t i[TRB_TOPK];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < TRB_TOPK; ++k) { s[k] = neg_inf<T>(); i[k] = INT64_MAX; }
    }
    __device__ __forceinline__ bool admits(T v, int64_t idx) const {
        return ranks_before(v, idx, s[TRB_TOPK - 1], i[TRB_TOPK - 1]);
    }
    __device__ __forceinline__ void push(T v, int64_t idx) {
        if (!admits(v, idx)) return;
        s[TRB_TOPK - 1] = v;
        i[TRB_TOPK - 1] = idx;
#pragma unroll
        for (int k = TRB_TOPK - 1; k > 0; --k) {
            if (ranks_before(s[k], i[k], s[k - 1], i[k - 1])) {
                T ts = s[k]; s[k] = s[k - 1]; s[k - 1] = ts;
                int64_t ti = i[k]; i[k] = i[k - 1]; i[k - 1] = ti;
            }
        }
    }
};
using TopK = TopKT<float>;
