// fp32 (FFMA) retrieval path: the parity path of the north star.  Every similarity is
// accumulated as acc = fmaf(q[k], g[k], acc), k ascending from acc = 0, in all three kernels
// that compute similarities, so thresholds and streamed values are bit-identical.
//
// Replaces lib/data/metrics/evaluation.py:117-120 (normalise, text @ image.T) and
// evaluation.py:11-37 (rank) of the reference without writing the [Q,G] matrix.
#include "common.cuh"
#include "sgemm.cuh"

// ---------------------------------------------------------------------------
// L2 row normalisation (F.normalize(p=2, dim=1, eps=1e-12))
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                float* __restrict__ inv_norm, int64_t rows,
                                                                int64_t dim, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * dim;
    float ss = 0.f;
    for (int64_t k = lane; k < dim; k += 32) { float v = xr[k]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), eps);
    if (y) {
        float* yr = y + row * dim;
        for (int64_t k = lane; k < dim; k += 32) yr[k] = __fdiv_rn(xr[k], nrm);
    }
    if (inv_norm && lane == 0) inv_norm[row] = __fdiv_rn(1.0f, nrm);
}

extern "C" int trb_l2_normalize_rows_f32(const float* x, float* y, float* inv_norm, int64_t rows, int64_t dim,
                                         float eps, trb_stream_t stream) {
    TRB_REQUIRE(x && (y || inv_norm), "l2_normalize: null pointer");
    TRB_REQUIRE(rows >= 0 && dim > 0, "l2_normalize: bad shape rows=%lld dim=%lld", (long long)rows, (long long)dim);
    if (rows == 0) return 0;
    const int wpb = 8;
    l2_normalize_rows_kernel<<<(unsigned)trb_ceil_div(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x, y, inv_norm, rows, dim, eps);
    TRB_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------
// thresholds: one thread per (query, relevant item) slot, sequential-k FFMA
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) thresholds_f32_kernel(const float* __restrict__ qn, const float* __restrict__ gn,
                                                             const int64_t* __restrict__ rel_ptr,
                                                             const int64_t* __restrict__ rel_row,
                                                             float* __restrict__ thr, int64_t Q, int64_t D) {
    // one warp walks the slots of one query so that the query row stays in L1
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int64_t lo = rel_ptr[q], hi = rel_ptr[q + 1];
    const float4* qr = reinterpret_cast<const float4*>(qn + q * D);
    for (int64_t slot = lo + lane; slot < hi; slot += 32) {
        const int64_t g = rel_row[slot];
        if (g < 0) continue;
        const float4* gr = reinterpret_cast<const float4*>(gn + g * D);
        float acc = 0.f;
        for (int64_t k4 = 0; k4 < D / 4; ++k4) {
            const float4 a = qr[k4], b = gr[k4];
            acc = fmaf(a.x, b.x, acc);
            acc = fmaf(a.y, b.y, acc);
            acc = fmaf(a.z, b.z, acc);
            acc = fmaf(a.w, b.w, acc);
        }
        thr[slot] = acc;
    }
}

extern "C" int trb_retrieval_thresholds_f32(const float* qn, const float* gn, const int64_t* rel_ptr,
                                            const int64_t* rel_row, float* thr, int64_t Q, int64_t D,
                                            trb_stream_t stream) {
    TRB_REQUIRE(qn && gn && rel_ptr && rel_row && thr, "thresholds_f32: null pointer");
    TRB_REQUIRE(Q >= 0 && D > 0 && D % 16 == 0, "thresholds_f32: D=%lld must be a positive multiple of 16", (long long)D);
    TRB_REQUIRE(trb_aligned16(qn) && trb_aligned16(gn), "thresholds_f32: operands must be 16-byte aligned");
    if (Q == 0) return 0;
    thresholds_f32_kernel<<<(unsigned)trb_ceil_div(Q, 4), 128, 0, (cudaStream_t)stream>>>(qn, gn, rel_ptr, rel_row, thr, Q, D);
    TRB_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------
// stream: 128x128 similarity tile per step, never written to HBM
// ---------------------------------------------------------------------------
namespace {
constexpr int BM = 128, BN = 128, BK = 16, S_LD = BN + 1, RREG = 8;
constexpr int STREAM_SMEM = (BK * BM + BK * BN + BM * S_LD) * (int)sizeof(float);
}

__global__ void __launch_bounds__(256, 1)
stream_f32_kernel(const float* __restrict__ qn, const float* __restrict__ gn, int64_t Q, int64_t G, int64_t D,
                  int64_t g_base, const int64_t* __restrict__ rel_ptr, const float* __restrict__ thr,
                  const int64_t* __restrict__ thr_gidx, int nsplit, float* __restrict__ cand_sim,
                  int64_t* __restrict__ cand_idx, int32_t* __restrict__ cnt) {
    extern __shared__ float smem[];
    float* As = smem;                 // [BK][BM]
    float* Bs = As + BK * BM;         // [BK][BN]
    float* S = Bs + BK * BN;          // [BM][S_LD]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t q0 = (int64_t)blockIdx.x * BM;
    const int split = blockIdx.y;
    const int64_t ntiles = (G + BN - 1) / BN;
    const int64_t t_lo = ntiles * split / nsplit, t_hi = ntiles * (split + 1) / nsplit;

    // epilogue roles: threads [0,128) keep the top-10 of row tid; threads [128,256) count for row tid-128
    const bool is_top = tid < BM;
    const int erow = is_top ? tid : tid - BM;
    const int64_t eq = q0 + erow;
    TopK top;
    top.init();
    int64_t s_lo = 0, s_hi = 0;
    float rthr[RREG];
    int64_t ridx[RREG];
    int rcnt[RREG];
#pragma unroll
    for (int r = 0; r < RREG; ++r) { rthr[r] = CUDART_INF_F; ridx[r] = -1; rcnt[r] = 0; }
    if (!is_top && rel_ptr != nullptr && eq < Q) {
        s_lo = rel_ptr[eq];
        s_hi = rel_ptr[eq + 1];
#pragma unroll
        for (int r = 0; r < RREG; ++r)
            if (s_lo + r < s_hi) { rthr[r] = thr[s_lo + r]; ridx[r] = thr_gidx[s_lo + r]; }
    }

    for (int64_t t = t_lo; t < t_hi; ++t) {
        const int64_t n0 = t * BN;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        for (int64_t k0 = 0; k0 < D; k0 += BK) {
            // 128 rows x 16 k = 512 float4 per operand; 256 threads -> 2 each
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int r = (tid >> 2) + p * 64, c4 = tid & 3;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (q0 + r < Q) a = *reinterpret_cast<const float4*>(qn + (q0 + r) * D + k0 + c4 * 4);
                if (n0 + r < G) b = *reinterpret_cast<const float4*>(gn + (n0 + r) * D + k0 + c4 * 4);
                As[(c4 * 4 + 0) * BM + r] = a.x; As[(c4 * 4 + 1) * BM + r] = a.y;
                As[(c4 * 4 + 2) * BM + r] = a.z; As[(c4 * 4 + 3) * BM + r] = a.w;
                Bs[(c4 * 4 + 0) * BN + r] = b.x; Bs[(c4 * 4 + 1) * BN + r] = b.y;
                Bs[(c4 * 4 + 2) * BN + r] = b.z; Bs[(c4 * 4 + 3) * BN + r] = b.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[8], b[8];
                const float4 a0 = *reinterpret_cast<const float4*>(As + kk * BM + ty * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(As + kk * BM + ty * 8 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * BN + tx * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(Bs + kk * BN + tx * 8 + 4);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) S[(ty * 8 + i) * S_LD + tx * 8 + j] = acc[i][j];
        __syncthreads();

        const int ncols = (int)min((int64_t)BN, G - n0);
        if (eq < Q) {
            const float* srow = S + erow * S_LD;
            if (is_top) {
                for (int c = 0; c < ncols; ++c) top.push(srow[c], g_base + n0 + c);
            } else if (s_hi > s_lo) {
                for (int c = 0; c < ncols; ++c) {
                    const float s = srow[c];
                    const int64_t gi = g_base + n0 + c;
#pragma unroll
                    for (int r = 0; r < RREG; ++r) rcnt[r] += ranks_before(s, gi, rthr[r], ridx[r]) ? 1 : 0;
                }
                for (int64_t slot = s_lo + RREG; slot < s_hi; ++slot) {   // rare: more than RREG relevant items
                    const float th = thr[slot];
                    const int64_t ti = thr_gidx[slot];
                    int c_ = 0;
                    for (int c = 0; c < ncols; ++c) c_ += ranks_before(srow[c], g_base + n0 + c, th, ti) ? 1 : 0;
                    if (c_) atomicAdd(cnt + slot, c_);
                }
            }
        }
        __syncthreads();
    }

    if (eq < Q) {
        if (is_top) {
            float* cs = cand_sim + (eq * nsplit + split) * TRB_TOPK;
            int64_t* ci = cand_idx + (eq * nsplit + split) * TRB_TOPK;
#pragma unroll
            for (int k = 0; k < TRB_TOPK; ++k) { cs[k] = top.s[k]; ci[k] = top.i[k]; }
        } else {
#pragma unroll
            for (int r = 0; r < RREG; ++r)
                if (s_lo + r < s_hi && rcnt[r]) atomicAdd(cnt + s_lo + r, rcnt[r]);
        }
    }
}

extern "C" int trb_retrieval_stream_f32(const float* qn, const float* gn, int64_t Q, int64_t G, int64_t D,
                                        int64_t g_base, const int64_t* rel_ptr, const float* thr,
                                        const int64_t* thr_gidx, int nsplit, float* cand_sim, int64_t* cand_idx,
                                        int32_t* cnt, trb_stream_t stream) {
    TRB_REQUIRE(qn && gn && cand_sim && cand_idx, "stream_f32: null pointer");
    TRB_REQUIRE(Q >= 0 && G >= 0 && D > 0 && D % 16 == 0, "stream_f32: D=%lld must be a positive multiple of 16", (long long)D);
    TRB_REQUIRE(nsplit >= 1 && nsplit <= 65535, "stream_f32: nsplit=%d out of range", nsplit);
    TRB_REQUIRE((rel_ptr == nullptr) == (thr == nullptr) && (thr == nullptr) == (thr_gidx == nullptr) &&
                    (thr == nullptr) == (cnt == nullptr),
                "stream_f32: rel_ptr, thr, thr_gidx and cnt must be given together");
    TRB_REQUIRE(trb_aligned16(qn) && trb_aligned16(gn), "stream_f32: operands must be 16-byte aligned");
    if (Q == 0) return 0;
    static TrbDeviceOnce attr_set;
    if (trb_first_on_device(attr_set))
        TRB_CUDA_OK(cudaFuncSetAttribute(stream_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
    dim3 grid((unsigned)trb_ceil_div(Q, BM), (unsigned)nsplit);
    stream_f32_kernel<<<grid, 256, STREAM_SMEM, (cudaStream_t)stream>>>(qn, gn, Q, G, D, g_base, rel_ptr, thr, thr_gidx,
                                                                       nsplit, cand_sim, cand_idx, cnt);
    TRB_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------
// rank from materialised scores: one CTA per query row.  The score of (q, g) comes from a loader:
//   SimLoader    : the fp32 similarity matrix itself                    (rank(similarity, ...), evaluation.py:11)
//   RerankLoader : float64  alpha * Jaccard(top-n(q), top-n(g)) + sim   (k-reciprocal re-ranking, evaluation.py:40-65,
//                  151-156: "rvn_mat + similarity" is float64 because jaccard_mat comes from np.zeros)
// ---------------------------------------------------------------------------
namespace {
constexpr int RS_THREADS = 256, RS_SLOTS = 64;

struct SimLoader {
    using T = float;
    const float* sim;
    int64_t row_stride, col_stride;
    const float* row;
    __device__ __forceinline__ void init(int64_t q) { row = sim + q * row_stride; }
    __device__ __forceinline__ float operator()(int64_t g) const { return row[g * col_stride]; }
};

struct F64Loader {           // a materialised float64 score matrix (re-ranked scores read back from inference_data.npz)
    using T = double;
    const double* m;
    int64_t row_stride, col_stride;
    const double* row;
    __device__ __forceinline__ void init(int64_t q) { row = m + q * row_stride; }
    __device__ __forceinline__ double operator()(int64_t g) const { return row[g * col_stride]; }
};

struct RerankLoader {
    using T = double;
    const float* sim;
    int64_t row_stride, col_stride;
    const int64_t* q_nn;     // [Q, n] neighbours of every query among the gallery
    const int64_t* g_nn;     // [G, n] neighbours of every gallery item among the gallery
    int n;
    double alpha;
    const float* row;
    const int64_t* qn;
    __device__ __forceinline__ void init(int64_t q) { row = sim + q * row_stride; qn = q_nn + q * n; }
    __device__ __forceinline__ double operator()(int64_t g) const {
        int inter = 0;
        for (int a = 0; a < n; ++a) {
            const int64_t x = qn[a];
            for (int b = 0; b < n; ++b) inter += (g_nn[g * n + b] == x) ? 1 : 0;
        }
        const double jac = (double)inter / (double)(2 * n - inter);        // |A & B| / |A | B| with |A| = |B| = n
        return alpha * jac + (double)row[g * col_stride];
    }
};
}  // namespace

template <class Loader>
__global__ void __launch_bounds__(RS_THREADS)
rank_scores_kernel(Loader ld, int64_t Q, int64_t G, const int64_t* __restrict__ rel_ptr, const int64_t* __restrict__ rel_col,
                   float* __restrict__ cand_sim, int64_t* __restrict__ cand_idx, int32_t* __restrict__ cnt) {
    using T = typename Loader::T;
    __shared__ T p_s[RS_THREADS * TRB_TOPK];
    __shared__ int64_t p_i[RS_THREADS * TRB_TOPK];
    __shared__ T th_s[RS_SLOTS];
    __shared__ int64_t th_i[RS_SLOTS];
    __shared__ int th_c[RS_SLOTS];
    __shared__ T w_s[RS_THREADS / 32];
    __shared__ int64_t w_i[RS_THREADS / 32];
    __shared__ int w_p[RS_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t q = blockIdx.x;
    ld.init(q);

    TopKT<T> top;
    top.init();
    for (int64_t g = tid; g < G; g += RS_THREADS) top.push(ld(g), g);

    // ranks of the relevant items, RS_SLOTS thresholds at a time
    if (rel_ptr != nullptr) {
        const int64_t s_lo = rel_ptr[q], s_hi = rel_ptr[q + 1];
        for (int64_t base = s_lo; base < s_hi; base += RS_SLOTS) {
            const int n = (int)min((int64_t)RS_SLOTS, s_hi - base);
            __syncthreads();
            if (tid < n) {
                const int64_t g = rel_col[base + tid];
                th_i[tid] = g;
                th_s[tid] = ld(g);
                th_c[tid] = 0;
            }
            __syncthreads();
            for (int r0 = 0; r0 < n; r0 += 8) {          // 8 thresholds per sweep: one score evaluation serves all of them
                int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int64_t g = tid; g < G; g += RS_THREADS) {
                    const T s = ld(g);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (r0 + j < n) c[j] += ranks_before(s, g, th_s[r0 + j], th_i[r0 + j]) ? 1 : 0;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int t = warp_sum_i(c[j]);
                    if (lane == 0 && t && r0 + j < n) atomicAdd(&th_c[r0 + j], t);
                }
            }
            __syncthreads();
            if (tid < n) cnt[base + tid] = th_c[tid];
        }
    }

    // merge the per-thread lists: 10 rounds of block arg-best over the pooled candidates
#pragma unroll
    for (int k = 0; k < TRB_TOPK; ++k) { p_s[tid * TRB_TOPK + k] = top.s[k]; p_i[tid * TRB_TOPK + k] = top.i[k]; }
    __syncthreads();
    for (int round = 0; round < TRB_TOPK; ++round) {
        T bs = neg_inf<T>();
        int64_t bi = INT64_MAX;
        int bp = -1;
        for (int p = tid; p < RS_THREADS * TRB_TOPK; p += RS_THREADS) {
            if (ranks_before(p_s[p], p_i[p], bs, bi)) { bs = p_s[p]; bi = p_i[p]; bp = p; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (ranks_before(os, oi, bs, bi)) { bs = os; bi = oi; bp = op; }
        }
        if (lane == 0) { w_s[wid] = bs; w_i[wid] = bi; w_p[wid] = bp; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < RS_THREADS / 32; ++w)
                if (ranks_before(w_s[w], w_i[w], bs, bi)) { bs = w_s[w]; bi = w_i[w]; bp = w_p[w]; }
            cand_sim[q * TRB_TOPK + round] = (float)bs;
            cand_idx[q * TRB_TOPK + round] = bi;
            if (bp >= 0) { p_s[bp] = neg_inf<T>(); p_i[bp] = INT64_MAX; }
        }
        __syncthreads();
    }
}

extern "C" int trb_rank_similarity_f32(const float* sim, int64_t row_stride, int64_t col_stride, int64_t Q, int64_t G,
                                       const int64_t* rel_ptr, const int64_t* rel_col, float* cand_sim,
                                       int64_t* cand_idx, int32_t* cnt, trb_stream_t stream) {
    TRB_REQUIRE(sim && cand_sim && cand_idx, "rank_similarity: null pointer");
    TRB_REQUIRE(Q >= 0 && G >= 0, "rank_similarity: bad shape");
    TRB_REQUIRE((rel_ptr == nullptr) == (rel_col == nullptr) && (rel_ptr == nullptr) == (cnt == nullptr),
                "rank_similarity: rel_ptr, rel_col and cnt must be given together");
    if (Q == 0) return 0;
    SimLoader ld{sim, row_stride, col_stride, nullptr};
    rank_scores_kernel<SimLoader><<<(unsigned)Q, RS_THREADS, 0, (cudaStream_t)stream>>>(ld, Q, G, rel_ptr, rel_col, cand_sim, cand_idx, cnt);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_rank_scores_f64(const double* scores, int64_t row_stride, int64_t col_stride, int64_t Q, int64_t G,
                                   const int64_t* rel_ptr, const int64_t* rel_col, float* cand_sim, int64_t* cand_idx,
                                   int32_t* cnt, trb_stream_t stream) {
    TRB_REQUIRE(scores && cand_sim && cand_idx, "rank_scores_f64: null pointer");
    TRB_REQUIRE(Q >= 0 && G >= 0, "rank_scores_f64: bad shape");
    TRB_REQUIRE((rel_ptr == nullptr) == (rel_col == nullptr) && (rel_ptr == nullptr) == (cnt == nullptr),
                "rank_scores_f64: rel_ptr, rel_col and cnt must be given together");
    if (Q == 0) return 0;
    F64Loader ld{scores, row_stride, col_stride, nullptr};
    rank_scores_kernel<F64Loader><<<(unsigned)Q, RS_THREADS, 0, (cudaStream_t)stream>>>(ld, Q, G, rel_ptr, rel_col, cand_sim, cand_idx, cnt);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_rank_rerank_f64(const float* sim, int64_t row_stride, int64_t col_stride, int64_t Q, int64_t G,
                                   const int64_t* q_nn, const int64_t* g_nn, int n_neighbors, double alpha,
                                   const int64_t* rel_ptr, const int64_t* rel_col, float* cand_sim, int64_t* cand_idx,
                                   int32_t* cnt, trb_stream_t stream) {
    TRB_REQUIRE(sim && q_nn && g_nn && cand_sim && cand_idx, "rank_rerank: null pointer");
    TRB_REQUIRE(Q >= 0 && G >= 0 && n_neighbors >= 1 && n_neighbors <= TRB_TOPK, "rank_rerank: bad shape (1 <= neighbours <= 10)");
    TRB_REQUIRE((rel_ptr == nullptr) == (rel_col == nullptr) && (rel_ptr == nullptr) == (cnt == nullptr),
                "rank_rerank: rel_ptr, rel_col and cnt must be given together");
    if (Q == 0) return 0;
    RerankLoader ld{sim, row_stride, col_stride, q_nn, g_nn, n_neighbors, alpha, nullptr, nullptr};
    rank_scores_kernel<RerankLoader><<<(unsigned)Q, RS_THREADS, 0, (cudaStream_t)stream>>>(ld, Q, G, rel_ptr, rel_col, cand_sim, cand_idx, cnt);
    TRB_LAUNCH_OK();
    return 0;
}

// alpha * Jaccard matrix in float64, materialised (the rvn_mat / rtn_mat arrays of inference_data.npz, evaluation.py:126-142)
__global__ void __launch_bounds__(256)
jaccard_f64_kernel(const int64_t* __restrict__ q_nn, const int64_t* __restrict__ g_nn, int n, double alpha, double* __restrict__ out,
                   int64_t Q, int64_t G) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Q * G) return;
    const int64_t q = e / G, g = e % G;
    int inter = 0;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) inter += (q_nn[q * n + a] == g_nn[g * n + b]) ? 1 : 0;
    out[e] = alpha * ((double)inter / (double)(2 * n - inter));
}

extern "C" int trb_jaccard_f64(const int64_t* q_nn, const int64_t* g_nn, int n_neighbors, double alpha, double* out, int64_t Q,
                               int64_t G, trb_stream_t stream) {
    TRB_REQUIRE(q_nn && g_nn && out, "jaccard: null pointer");
    TRB_REQUIRE(Q >= 0 && G >= 0 && n_neighbors >= 1, "jaccard: bad shape");
    if (Q * G == 0) return 0;
    jaccard_f64_kernel<<<(unsigned)trb_ceil_div(Q * G, 256), 256, 0, (cudaStream_t)stream>>>(q_nn, g_nn, n_neighbors, alpha, out, Q, G);
    TRB_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------
// finish: merge candidate lists, hit ranks, AP.  One warp per query.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
retrieval_finish_kernel(const float* __restrict__ cand_sim, const int64_t* __restrict__ cand_idx, int nlists, int64_t Q,
                        const int64_t* __restrict__ q_pids, const int64_t* __restrict__ g_pids, int64_t G_total,
                        const int64_t* __restrict__ rel_ptr, const int32_t* __restrict__ cnt, float* __restrict__ top_sim,
                        int64_t* __restrict__ top_idx, int32_t* __restrict__ first_hit, int32_t* __restrict__ hit_ranks,
                        float* __restrict__ ap) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int n = nlists * TRB_TOPK;
    const float* cs = cand_sim + q * n;
    const int64_t* ci = cand_idx + q * n;

    int first_in_top = INT32_MAX;
    const int64_t qpid = (q_pids != nullptr) ? q_pids[q] : -1;
    constexpr int CPL = 8;                       // candidates per lane kept in registers (n <= 256)
    if (nlists == 1) {
        // already best-first (possibly ordered by float64 scores that the fp32 copies cannot distinguish): pass through
        if (lane < TRB_TOPK) {
            int64_t i = ci[lane];
            if (i == INT64_MAX) i = -1;
            top_sim[q * TRB_TOPK + lane] = cs[lane];
            top_idx[q * TRB_TOPK + lane] = i;
        }
        if (g_pids != nullptr) {
            for (int r = 0; r < TRB_TOPK; ++r) {
                const int64_t i = ci[r];
                if (i >= 0 && i < G_total && g_pids[i] == qpid) { first_in_top = r; break; }
            }
        }
    } else if (n <= 32 * CPL) {
        float rs[CPL];
        int64_t ri[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int p = lane + 32 * c;
            rs[c] = -CUDART_INF_F;
            ri[c] = INT64_MAX;
            if (p < n) {
                rs[c] = cs[p];
                ri[c] = ci[p];
                if (ri[c] < 0) ri[c] = INT64_MAX;            // padding of an already-merged list
            }
        }
        for (int round = 0; round < TRB_TOPK; ++round) {
            float bs = -CUDART_INF_F;
            int64_t bi = INT64_MAX;
            int bc = -1;
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                if (ranks_before(rs[c], ri[c], bs, bi)) { bs = rs[c]; bi = ri[c]; bc = c; }
            float ws = bs;
            int64_t wi = bi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, ws, o);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, wi, o);
                if (ranks_before(os, oi, ws, wi)) { ws = os; wi = oi; }
            }
            if (bc >= 0 && bs == ws && bi == wi) {           // this lane owned the winner: consume it
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    if (c == bc) { rs[c] = -CUDART_INF_F; ri[c] = INT64_MAX; }
            }
            if (lane == 0) {
                top_sim[q * TRB_TOPK + round] = ws;
                top_idx[q * TRB_TOPK + round] = (wi == INT64_MAX) ? -1 : wi;
                if (first_in_top == INT32_MAX && g_pids != nullptr && wi != INT64_MAX && wi >= 0 && wi < G_total && g_pids[wi] == qpid)
                    first_in_top = round;
            }
        }
    } else {
        float last_s = CUDART_INF_F;
        int64_t last_i = -1;
        for (int round = 0; round < TRB_TOPK; ++round) {
            float bs = -CUDART_INF_F;
            int64_t bi = INT64_MAX;
            for (int p = lane; p < n; p += 32) {
                const float s = cs[p];
                int64_t i = ci[p];
                if (i < 0) i = INT64_MAX;
                // strictly after the previous winner, and better than the running best
                if (ranks_before(last_s, last_i, s, i) && ranks_before(s, i, bs, bi)) { bs = s; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, bs, o);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ranks_before(os, oi, bs, bi)) { bs = os; bi = oi; }
            }
            if (lane == 0) {
                top_sim[q * TRB_TOPK + round] = bs;
                top_idx[q * TRB_TOPK + round] = (bi == INT64_MAX) ? -1 : bi;
                if (first_in_top == INT32_MAX && g_pids != nullptr && bi != INT64_MAX && bi >= 0 && bi < G_total && g_pids[bi] == qpid)
                    first_in_top = round;
            }
            last_s = bs;
            last_i = bi;
        }
    }

    if (lane != 0) return;
    if (rel_ptr == nullptr) {
        if (first_hit) first_hit[q] = first_in_top;
        return;
    }
    const int64_t lo = rel_ptr[q], hi = rel_ptr[q + 1];
    const int64_t R = hi - lo;
    // ascending insertion sort of the ranks (R is the number of gallery images of one identity)
    float sum = 0.f;
    int best = INT32_MAX;
    if (hit_ranks != nullptr) {
        for (int64_t a = 0; a < R; ++a) {
            const int32_t v = cnt[lo + a];
            int64_t b = a;
            while (b > 0 && hit_ranks[lo + b - 1] > v) { hit_ranks[lo + b] = hit_ranks[lo + b - 1]; --b; }
            hit_ranks[lo + b] = v;
        }
        for (int64_t j = 0; j < R; ++j) {
            const int32_t r = hit_ranks[lo + j];
            sum = __fadd_rn(sum, __fdiv_rn((float)(j + 1), (float)(r + 1)));
        }
        if (R > 0) best = hit_ranks[lo];
    } else {
        // no scratch for sorted ranks: selection by repeated minimum (O(R^2), R is small)
        int32_t prev = -1;
        for (int64_t j = 0; j < R; ++j) {
            int32_t m = INT32_MAX;
            for (int64_t a = 0; a < R; ++a) { const int32_t v = cnt[lo + a]; if (v > prev && v < m) m = v; }
            if (j == 0) best = m;
            sum = __fadd_rn(sum, __fdiv_rn((float)(j + 1), (float)(m + 1)));
            prev = m;
        }
    }
    if (first_hit) first_hit[q] = best;
    if (ap) ap[q] = __fdiv_rn(sum, (float)R);   // 0/0 -> NaN exactly like the reference
}

extern "C" int trb_retrieval_finish(const float* cand_sim, const int64_t* cand_idx, int nlists, int64_t Q,
                                    const int64_t* q_pids, const int64_t* g_pids, int64_t G_total,
                                    const int64_t* rel_ptr, const int32_t* cnt, float* top_sim, int64_t* top_idx,
                                    int32_t* first_hit, int32_t* hit_ranks, float* ap, trb_stream_t stream) {
    TRB_REQUIRE(cand_sim && cand_idx && top_sim && top_idx, "retrieval_finish: null pointer");
    TRB_REQUIRE(rel_ptr != nullptr || first_hit == nullptr || (q_pids && g_pids),
                "retrieval_finish: top-k-only first hits need q_pids and g_pids");
    TRB_REQUIRE(nlists >= 1 && Q >= 0, "retrieval_finish: bad shape");
    TRB_REQUIRE((rel_ptr == nullptr) == (cnt == nullptr), "retrieval_finish: rel_ptr and cnt must be given together");
    if (Q == 0) return 0;
    retrieval_finish_kernel<<<(unsigned)trb_ceil_div(Q, 4), 128, 0, (cudaStream_t)stream>>>(
        cand_sim, cand_idx, nlists, Q, q_pids, g_pids, G_total, rel_ptr, cnt, top_sim, top_idx, first_hit, hit_ranks, ap);
    TRB_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------
// metrics: CMC@k counts and mean AP, fixed reduction order (one CTA)
// ---------------------------------------------------------------------------
struct TopkCuts { int32_t k[8]; int n; };

__global__ void __launch_bounds__(1024)
retrieval_metrics_kernel(const int32_t* __restrict__ first_hit, const float* __restrict__ ap, int64_t Q, TopkCuts cuts,
                         float* __restrict__ cmc_out, float* __restrict__ map_out) {
    __shared__ double sd[1024];
    __shared__ int si[8][32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double s = 0.0;
    // fixed assignment q = tid + 1024*j, four independent loads in flight per thread
    for (int64_t q0 = tid; q0 < Q; q0 += 4096) {
        int32_t fh[4];
        float a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t q = q0 + 1024 * u;
            fh[u] = q < Q ? first_hit[q] : INT32_MAX;
            a[u] = (ap && q < Q) ? ap[q] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] += (i < cuts.n && fh[u] < cuts.k[i]) ? 1 : 0;
            s += (double)a[u];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int v = warp_sum_i(c[i]);
        if (lane == 0) si[i][wid] = v;
    }
    sd[tid] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) sd[tid] += sd[tid + o];
        __syncthreads();
    }
    if (tid < cuts.n) {
        int tot = 0;
        for (int w = 0; w < 32; ++w) tot += si[tid][w];
        cmc_out[tid] = __fmul_rn(__fdiv_rn((float)tot, (float)Q), 100.0f);
    }
    if (tid == 0 && map_out && ap) *map_out = __fmul_rn((float)(sd[0] / (double)Q), 100.0f);
}

extern "C" int trb_retrieval_metrics(const int32_t* first_hit, const float* ap, int64_t Q, const int32_t* host_topk,
                                     int n_topk, float* cmc_out, float* map_out, trb_stream_t stream) {
    TRB_REQUIRE(first_hit && host_topk && cmc_out, "retrieval_metrics: null pointer");
    TRB_REQUIRE(n_topk >= 1 && n_topk <= 8, "retrieval_metrics: n_topk=%d must be in [1,8]", n_topk);
    TRB_REQUIRE(Q > 0, "retrieval_metrics: Q must be positive");
    TopkCuts cuts;
    cuts.n = n_topk;
    for (int i = 0; i < 8; ++i) cuts.k[i] = i < n_topk ? host_topk[i] : 0;
    retrieval_metrics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(first_hit, ap, Q, cuts, cmc_out, map_out);
    TRB_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------
// materialised similarity (compatibility: save_data npz, re-ranking, rank() callers)
// ---------------------------------------------------------------------------
extern "C" int trb_similarity_f32(const float* qn, const float* gn, float* sim, int64_t Q, int64_t G, int64_t D,
                                  trb_stream_t stream) {
    TRB_REQUIRE(qn && gn && sim, "similarity_f32: null pointer");
    TRB_REQUIRE(Q >= 0 && G >= 0 && D > 0, "similarity_f32: bad shape");
    TRB_REQUIRE(Q < (1LL << 31) && G < (1LL << 31) && D < (1LL << 31), "similarity_f32: dimension exceeds int32");
    if (Q == 0 || G == 0) return 0;
    GemmArgs g{qn, D, 1, gn, 1, D, sim, G, 0, (int)Q, (int)G, (int)D, nullptr, nullptr, 1};
    return launch_gemm(g, (cudaStream_t)stream);
}
