// HBM-bound elementwise kernels of the MoCo step: momentum (EMA) update of the key encoders,
// queue enqueue, and the tiny gradient-combination helpers.  128-bit accesses, grid sized in
// multiples of the SM count, streaming cache hints (each byte is touched once per step).
//
// Reference: lib/models/embeddings/moco_head/head.py:73-94 (_momentum_update_key_encoder) and
// :96-109 (_dequeue_and_enqueue).
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_rw(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// the reference's arithmetic: two rounded products, one rounded sum (no FMA contraction)
__device__ __forceinline__ float ema1(float k, float q, float m, float om) {
    return __fadd_rn(__fmul_rn(k, m), __fmul_rn(q, om));
}
__device__ __forceinline__ float4 ema4(const float4& k, const float4& q, float m, float om) {
    return make_float4(ema1(k.x, q.x, m, om), ema1(k.y, q.y, m, om), ema1(k.z, q.z, m, om), ema1(k.w, q.w, m, om));
}

constexpr int EMA_THREADS = 256, EMA_UNROLL = 4;

// flat arena: grid-stride over float4 with 4 independent 128-bit loads per operand in flight
__global__ void __launch_bounds__(EMA_THREADS)
ema_flat_kernel(float* __restrict__ pk, const float* __restrict__ pq, int64_t n, float m, float om) {
    const int64_t n4 = n >> 2;
    float4* k4 = reinterpret_cast<float4*>(pk);
    const float4* q4 = reinterpret_cast<const float4*>(pq);
    const int64_t stride = (int64_t)gridDim.x * EMA_THREADS;
    int64_t i = (int64_t)blockIdx.x * EMA_THREADS + threadIdx.x;
    for (; i + (EMA_UNROLL - 1) * stride < n4; i += EMA_UNROLL * stride) {
        float4 kv[EMA_UNROLL], qv[EMA_UNROLL];
#pragma unroll
        for (int u = 0; u < EMA_UNROLL; ++u) { kv[u] = ld_rw(k4 + i + u * stride); qv[u] = ld_stream(q4 + i + u * stride); }
#pragma unroll
        for (int u = 0; u < EMA_UNROLL; ++u) st_stream(k4 + i + u * stride, ema4(kv[u], qv[u], m, om));
    }
    for (; i < n4; i += stride) st_stream(k4 + i, ema4(ld_rw(k4 + i), ld_stream(q4 + i), m, om));
    // scalar tail (n not a multiple of 4)
    const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * EMA_THREADS + threadIdx.x;
    if (t < n) pk[t] = ema1(pk[t], pq[t], m, om);
}

// multi-tensor: blockIdx.x walks the chunk table; a chunk is at most max_chunk elements
__global__ void __launch_bounds__(EMA_THREADS)
ema_chunks_kernel(const trb_ema_chunk* __restrict__ chunks, int64_t nchunks, float m, float om) {
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const trb_ema_chunk ch = chunks[c];
        float* pk = ch.k;
        const float* pq = ch.q;
        const int64_t n = ch.n;
        const bool vec = (((uintptr_t)pk | (uintptr_t)pq) & 15u) == 0;
        if (vec) {
            const int64_t n4 = n >> 2;
            float4* k4 = reinterpret_cast<float4*>(pk);
            const float4* q4 = reinterpret_cast<const float4*>(pq);
            int64_t i = threadIdx.x;
            for (; i + (EMA_UNROLL - 1) * EMA_THREADS < n4; i += EMA_UNROLL * EMA_THREADS) {
                float4 kv[EMA_UNROLL], qv[EMA_UNROLL];
#pragma unroll
                for (int u = 0; u < EMA_UNROLL; ++u) { kv[u] = ld_rw(k4 + i + u * EMA_THREADS); qv[u] = ld_stream(q4 + i + u * EMA_THREADS); }
#pragma unroll
                for (int u = 0; u < EMA_UNROLL; ++u) st_stream(k4 + i + u * EMA_THREADS, ema4(kv[u], qv[u], m, om));
            }
            for (; i < n4; i += EMA_THREADS) st_stream(k4 + i, ema4(ld_rw(k4 + i), ld_stream(q4 + i), m, om));
            for (int64_t t = (n4 << 2) + threadIdx.x; t < n; t += EMA_THREADS) pk[t] = ema1(pk[t], pq[t], m, om);
        } else {
            for (int64_t t = threadIdx.x; t < n; t += EMA_THREADS) pk[t] = ema1(pk[t], pq[t], m, om);
        }
    }
}

// enqueue: keys [N,D] -> queue[:, ptr:ptr+N] (queue is [D,K], K contiguous): 32x32 smem transpose,
// both queues in one launch (blockIdx.z), ids by z==0,y==0 CTAs.  The pointer is advanced by a
// second one-thread launch so that no CTA can observe the new value (no cross-CTA ordering needed).
__global__ void __launch_bounds__(256)
enqueue_kernel(float* __restrict__ v_queue, float* __restrict__ t_queue, int64_t* __restrict__ id_queue,
               int64_t* __restrict__ queue_ptr, const float* __restrict__ v_keys, const float* __restrict__ t_keys,
               const int64_t* __restrict__ ids, int N, int D, int K) {
    __shared__ float tile[32][33];
    const int64_t ptr = *queue_ptr;
    const float* keys = blockIdx.z ? t_keys : v_keys;
    float* queue = blockIdx.z ? t_queue : v_queue;
    const int n0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, d = d0 + tx;
        tile[r][tx] = (n < N && d < D) ? keys[(int64_t)n * D + d] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r, n = n0 + tx;
        // a pointer outside [0, K-N] (a checkpoint written with another batch size: the reference's slice assignment raises
        // there, head.py:104) wraps around instead of running over the end of the row / of the allocation
        if (n < N && d < D) queue[(int64_t)d * K + (((ptr + n) % K) + K) % K] = tile[tx][r];
    }
    if (blockIdx.z == 0 && blockIdx.y == 0 && ty == 0 && n0 + tx < N) id_queue[(((ptr + n0 + tx) % K) + K) % K] = ids[n0 + tx];

}

__global__ void advance_queue_ptr_kernel(int64_t* __restrict__ queue_ptr, int N, int K) {
    *queue_ptr = (((*queue_ptr + N) % K) + K) % K;
}

__global__ void __launch_bounds__(256)
combine3_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                const float* __restrict__ g, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = 0.f;
    if (a) v = fmaf(g[0], a[i], v);
    if (b) v = fmaf(g[1], b[i], v);
    if (c) v = fmaf(g[2], c[i], v);
    out[i] = v;
}

__global__ void __launch_bounds__(256) scale_inplace_kernel(float* __restrict__ x, const float* __restrict__ g, int64_t n) {
    const float s = g[0];
    if (s == 1.0f) return;      // the trainer's case (trainer.py:82): the gradient is already final
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

// backward of the loss dict in one launch: per-loss gradients x upstream scalars (device pointers, NULL = 0) -> final gradients.
// items [0, 2*nd): embedding gradients (modality = item / nd); items [2*nd, 2*nd + dc): projection gradient.
__global__ void __launch_bounds__(256)
grad_combine_kernel(const float* __restrict__ d_inst, const float* __restrict__ d_nce, const float* __restrict__ d_ga,
                    const float* d_proj, const float* __restrict__ g_inst, const float* __restrict__ g_nce,
                    const float* __restrict__ g_ga, int separate_q, int64_t nd, int64_t dc, float* __restrict__ out_v,
                    float* __restrict__ out_t, float* __restrict__ out_vq, float* __restrict__ out_tq, float* out_proj,
                    int vec4) {
    const float gi = g_inst ? g_inst[0] : 0.f, gn = g_nce ? g_nce[0] : 0.f, gg = g_ga ? g_ga[0] : 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t e = tid; e < 2 * nd; e += stride) {
        const int64_t i = e < nd ? e : e - nd;
        float v = fmaf(gi, d_inst[e], 0.f);
        if (separate_q) {
            v = fmaf(gg, d_ga[e], v);
            (e < nd ? out_vq : out_tq)[i] = fmaf(gn, d_nce[e], 0.f);
        } else {
            v = fmaf(gn, d_nce[e], v);
            v = fmaf(gg, d_ga[e], v);
        }
        (e < nd ? out_v : out_t)[i] = v;
    }
    if (out_proj == nullptr) return;
    if (out_proj == d_proj && gi == 1.0f) return;      // in-place hand-over with upstream gradient 1 (trainer.py:82): no traffic
    if (vec4) {
        const float4* src = reinterpret_cast<const float4*>(d_proj);
        float4* dst = reinterpret_cast<float4*>(out_proj);
        for (int64_t i = tid; i < dc / 4; i += stride) {
            float4 x = src[i];
            if (gi != 1.0f) { x.x *= gi; x.y *= gi; x.z *= gi; x.w *= gi; }
            dst[i] = x;
        }
    } else {
        for (int64_t i = tid; i < dc; i += stride) out_proj[i] = (gi != 1.0f) ? d_proj[i] * gi : d_proj[i];
    }
}

}  // namespace

extern "C" int trb_moco_grad_combine(const float* d_inst, const float* d_nce, const float* d_ga, const float* d_proj,
                                     const float* g_inst, const float* g_nce, const float* g_ga, int separate_q, int64_t nd,
                                     int64_t dc, float* out_v, float* out_t, float* out_vq, float* out_tq, float* out_proj,
                                     trb_stream_t stream) {
    TRB_REQUIRE(d_inst && d_nce && d_ga && out_v && out_t, "grad_combine: null pointer");
    TRB_REQUIRE(!separate_q || (out_vq && out_tq), "grad_combine: separate_q needs out_vq / out_tq");
    TRB_REQUIRE(out_proj == nullptr || d_proj != nullptr, "grad_combine: out_proj needs d_proj");
    TRB_REQUIRE(nd >= 0 && dc >= 0, "grad_combine: negative size");
    if (nd == 0 && (dc == 0 || out_proj == nullptr)) return 0;
    const int vec4 = out_proj != nullptr && (dc % 4) == 0 && trb_aligned16(d_proj) && trb_aligned16(out_proj);
    const int64_t items = 2 * nd > (vec4 ? dc / 4 : dc) ? 2 * nd : (vec4 ? dc / 4 : dc);
    int64_t blocks = trb_ceil_div(items, 256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    grad_combine_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_inst, d_nce, d_ga, d_proj, g_inst, g_nce, g_ga,
                                                                             separate_q, nd, dc, out_v, out_t, out_vq, out_tq,
                                                                             out_proj, vec4);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_ema_update_f32(float* p_k, const float* p_q, int64_t n, float m, float one_minus_m, trb_stream_t stream) {
    TRB_REQUIRE(p_k && p_q, "ema_update: null pointer");
    TRB_REQUIRE(n >= 0, "ema_update: negative length");
    TRB_REQUIRE(trb_aligned16(p_k) && trb_aligned16(p_q), "ema_update: arenas must be 16-byte aligned");
    if (n == 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = trb_ceil_div(n >> 2, (int64_t)EMA_THREADS * EMA_UNROLL);
    int64_t cap = (int64_t)sms * 16;
    unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    ema_flat_kernel<<<grid, EMA_THREADS, 0, (cudaStream_t)stream>>>(p_k, p_q, n, m, one_minus_m);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_ema_update_chunks_f32(const trb_ema_chunk* chunks, int64_t nchunks, int64_t max_chunk, float m,
                                         float one_minus_m, trb_stream_t stream) {
    TRB_REQUIRE(chunks || nchunks == 0, "ema_update_chunks: null table");
    TRB_REQUIRE(nchunks >= 0 && max_chunk > 0, "ema_update_chunks: bad sizes");
    if (nchunks == 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t cap = (int64_t)sms * 16;
    unsigned grid = (unsigned)(nchunks > cap ? cap : nchunks);
    ema_chunks_kernel<<<grid, EMA_THREADS, 0, (cudaStream_t)stream>>>(chunks, nchunks, m, one_minus_m);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_enqueue(float* v_queue, float* t_queue, int64_t* id_queue, int64_t* queue_ptr, const float* v_keys,
                           const float* t_keys, const int64_t* ids, int32_t N, int32_t D, int32_t K, trb_stream_t stream) {
    TRB_REQUIRE(v_queue && t_queue && id_queue && queue_ptr && v_keys && t_keys && ids, "enqueue: null pointer");
    TRB_REQUIRE(N > 0 && D > 0 && K > 0, "enqueue: bad shape");
    TRB_REQUIRE(K % N == 0, "enqueue: K=%d must be a multiple of the batch size N=%d (head.py:101)", K, N);
    dim3 grid((N + 31) / 32, (D + 31) / 32, 2);
    enqueue_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(v_queue, t_queue, id_queue, queue_ptr, v_keys, t_keys, ids, N, D, K);
    TRB_LAUNCH_OK();
    advance_queue_ptr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(queue_ptr, N, K);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_combine3_f32(float* out, const float* a, const float* b, const float* c, const float* g, int64_t n,
                                trb_stream_t stream) {
    TRB_REQUIRE(out && g, "combine3: null pointer");
    if (n <= 0) return 0;
    combine3_kernel<<<(unsigned)trb_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(out, a, b, c, g, n);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_scale_inplace_f32(float* x, const float* g, int64_t n, trb_stream_t stream) {
    TRB_REQUIRE(x && g, "scale_inplace: null pointer");
    if (n <= 0) return 0;
    int64_t blocks = trb_ceil_div(n, 256 * 8);
    scale_inplace_kernel<<<(unsigned)(blocks > 4736 ? 4736 : blocks), 256, 0, (cudaStream_t)stream>>>(x, g, n);
    TRB_LAUNCH_OK();
    return 0;
}
