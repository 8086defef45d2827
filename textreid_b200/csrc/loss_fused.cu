// Fused bf16 MoCo loss step for sm_100a: the whole loss dict AND its gradients in one cooperative kernel.
//
// Math: SURVEY.md section 8a' (head.py:148-170, losses.py:6-62,102-128,206-217).  Work decomposition: one CTA per
//   * instance tile  : 128 classes of the projection W [D, C]          (ceil(C/128) CTAs)
//   * InfoNCE tile   : 128 slots of one modality's queue [D, K]        (2 * ceil(K/128) CTAs)
//   * align          : the N x N similarity of the batch               (1 CTA)
// A tile CTA streams its fp32 W / queue tile from HBM exactly once (rounded to bf16 into shared memory in the K/MN-major
// SWIZZLE_128B image tcgen05.mma reads, column norms taken on the way), keeps the 2N embedding rows resident (bulk copy of
// the packed image the prologue wrote), and runs
//     z    = E  . Wb          tcgen05.mma  A K-major,  B MN-major     -> TMEM          (logits never leave the SM)
//     softmax partial statistics per row -> global, GRID BARRIER, every CTA combines the statistics of all tiles
//     dz'  = (softmax - target) / N * colscale -> bf16 in shared memory (K-major for dE, MN-major for dW)
//     dE_p = dz' . Wb^T       tcgen05.mma  A K-major,  B K-major      -> partial [tile] in global (L2)
//     dWs  = E^T . dz'        tcgen05.mma  A MN-major, B MN-major     -> TMEM -> shared -> column-normalisation Jacobian -> dW
//     GRID BARRIER, fixed-order reduction of the dE partials by all CTAs (+ normalise-backward for InfoNCE), loss scalars.
// The same embedding image serves as K-major A (forward) and MN-major A (dW), the same W image as MN-major B (forward) and
// K-major B (dE): nothing is transposed or re-packed.  HBM traffic = W + queues read once, dW written once.
#include "tc_common.cuh"
#include "loss_fused.cuh"
#include <stdlib.h>

namespace {

using namespace tc;

constexpr int F_THREADS = 256;
constexpr int FIN_PARTS = 4;                        // final reduction: threads per 16-byte slot
constexpr int F_TILE = 128;                         // classes / queue slots per CTA
constexpr int F_E_BYTES = 8 * BLOCK_BYTES;          // 128 KiB  embedding rows: [2 row-blocks][<=4 k-chunks][16 KiB]
constexpr int F_DZ_BYTES = 2 * BLOCK_BYTES;         // 32 KiB   logit gradient of one 128-row block: [2 column chunks][16 KiB]
constexpr int F_WB_BYTES = 2 * 256 * 128;           // 64 KiB   W / queue tile: [2 column chunks][256 d][128 B]
constexpr int F_WB_CHUNK = 256 * 128;
constexpr int F_OFF_E = 0, F_OFF_DZ = F_E_BYTES, F_OFF_WB = F_OFF_DZ + F_DZ_BYTES, F_OFF_MISC = F_OFF_WB + F_WB_BYTES;
constexpr int F_MISC_BYTES = 2048;
constexpr int F_SMEM = F_OFF_MISC + F_MISC_BYTES + 1024;   // + alignment slack

struct ProArgs {
    const float *v_embed, *t_embed, *v_qraw, *t_qraw, *v_key, *t_key, *v_queue, *t_queue;
    float *v_key_n, *t_key_n, *E2, *en, *inv_e, *qn, *inv_q, *pos;
    uint8_t *Ep, *ENp, *QNp, *QUp;
    int normalize_keys, N, D, KC, K, T_k;      // N = the whole batch: ceil(N / 128) windows of 2 x 128 padded rows
    const float* W;    // projection [D, C] for the L2 prefetch (NULL: none)
    int64_t W_bytes;
};

struct FP {
    int N, D, K, C, KC, T_inst, T_k, n_inst, n_nce, n_ga, want_grad, reduce_losses, roles;
    // Row windows: a batch of more than 128 rows is processed in ceil(N / 128) windows of <= 128 rows per modality INSIDE the
    // kernel (tile_program); row-indexed arrays are [modality][N rows], NS (= N) the stride between the modalities, Nn (= N) the
    // batch size every mean is taken over.  The queue mask (head.py:148-157) always uses the ids of the whole batch.
    int NS, Nn;
    int nce_dual;                   // an InfoNCE CTA takes tile t of BOTH modalities, one after the other (n_nce = T_k CTAs): the
                                    // instance tiles run more than twice as long, and 86 + 2 x 32 tiles would not fit 148 SMs
    int fin_U;                      // 16-byte slots per work unit of the partial-tile reduction (layout of part_inst)
    int row_helpers;                // this many spare CTAs form the instance row losses (0: tile 0 does)
    int fin_early;                  // the partial reductions belong to the CTAs behind the instance tiles (see fused_loss_kernel)
    float T, eps, alpha, beta, sp, sn;
    const float* W;
    const float* queue[2];          // queue scored by modality m's queries: [0] = t_queue, [1] = v_queue  (head.py:162,168)
    const float* key_n[2];          // positive key of modality m: [0] = t_key_n, [1] = v_key_n            (head.py:160,166)
    const int64_t *labels, *id_queue;
    const uint8_t *Ep, *ENp, *QNp;
    const uint8_t* QUp;             // bf16 tile images of the two queues [modality][tile][64 KiB], written by the prologue
    const float *en, *qn, *inv_e, *inv_q, *pos;
    float *ls_inst, *ls_nce;                // per-tile softmax statistics [tile][256 rows]: log2 sum_c 2^(z2_c) over the tile's columns
    float2* zz_inst;                        // ... and (sum z, z_y) of the instance tiles (label smoothing / target logit)
    unsigned long long* dbg;                // optional phase timestamps [cta][16] (TRB_FUSED_DEBUG)
    float* ga_part;                         // align, batches above 128 rows: dq_t contribution of block (iw, jw), fp32 [128][Dp]
    uint4 *part_inst, *part_nce;            // partial dE tiles, bf16 x 8 per 16 bytes: InfoNCE [tile][128 rows][8-column chunk],
                                            // instance [unit][tile][U slots] (see PartialReducer)
    float *dpos, *rows_inst, *rows_nce, *rows_ga, *losses, *d_inst, *d_nce, *d_ga, *d_proj;
    unsigned* bar;                  // [0] instance statistics, [1] partial tiles written, [2], [3] InfoNCE statistics per modality,
                                    // [4] next work unit of the partial-tile reduction;
                                    // zeroed by the prologue launch
    // _dequeue_and_enqueue (head.py:96-109) folded into the kernel: the InfoNCE / align CTAs, which reach the second grid barrier
    // long before the instance tiles, write the normalised keys and ids into the queues; the pointer moves after the barrier
    float* enq_queue[2];            // [0] = v_queue <- v_key_n, [1] = t_queue <- t_key_n; NULL = no enqueue
    int64_t *enq_ids, *enq_ptr;
    float* dbg_logits;              // optional [256][128] fp32 logits of instance tile dbg_tile (trb_moco_loss_debug_logits)
    int dbg_tile;
};

// Programmatic dependent launch: the cooperative kernel is launched while the prologue still runs (its CTAs start as SMs free
// up) and streams its W tiles; everything the prologue writes is touched only after griddep_wait().  Both are no-ops for a
// launch without the attribute.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MN-major SWIZZLE_128B descriptor (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): the 64-element groups
// along M/N are `lbo` bytes apart, the 8-row groups along K 1024 bytes.
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return umma_idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// Grid-wide counters in two halves.  arrive: everything this CTA wrote so far is published (bar.sync, then one thread fences and
// bumps the counter); wait: one thread spins until `total` CTAs have arrived.  CTAs that only produce (the instance tiles at
// the second counter) arrive and move on; CTAs that only consume (the spare CTAs) wait without arriving.
__device__ __forceinline__ void grid_arrive(unsigned* ctr) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
    }
}
__device__ __forceinline__ void grid_wait(const unsigned* ctr, unsigned total) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= total) break;
            __nanosleep(20);
            if (clock64() - t0 > 4000000000LL) wait_timed_out(1);
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t wb_offset(int d, int c) {        // bf16 element (d, c) of the W / queue tile image
    return (uint32_t)((c >> 6) * F_WB_CHUNK + (d >> 3) * 1024 + (d & 7) * 128 + ((((c & 63) >> 3) ^ (d & 7)) << 4) + (c & 7) * 2);
}
__device__ __forceinline__ uint32_t dz_offset(int h, int n, int c16) {   // 16-byte chunk c16 of row n, column chunk h
    return (uint32_t)(h * BLOCK_BYTES + (n >> 3) * 1024 + (n & 7) * 128 + ((c16 ^ (n & 7)) << 4));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// Column order of warp w's rows (d = w mod 8): columns are rotated by `head` so that every warp-wide access of 32 consecutive
// floats starts on a 32-byte sector of the row-major [*, ld] matrix (rows are only 4-byte aligned when ld is odd); the one
// access that wraps around holds the two ragged tile edges.
__device__ __forceinline__ int sector_head(int w, int64_t ld, int c0) { return (8 - (int)(((int64_t)w * ld + c0) & 7)) & 7; }

// fp32 tile src[d, c0 + c] (d < nrows, c < 128) -> bf16 image in shared memory; per-column sums of squares -> red[8][128].
// Warp w takes rows d = w + 8 i, so (d & 7) == w and (d >> 3) == i: the swizzled column part of the address is a per-thread
// constant.  All loads of the tile are issued before the first use; `rot` staggers the row order between CTAs so that CTAs
// sweeping the same rows of a power-of-two-pitched matrix do not hit the same DRAM channels in lock step.
template <typename Mid>
__device__ __forceinline__ void load_tile_bf16(const float* __restrict__ src, int64_t ld, int nrows, int rows_pad, int c0,
                                               int ncols, int rot, uint8_t* wb, float* red, Mid&& mid) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = sector_head(w, ld, c0);
    const int groups = rows_pad >> 3;                      // 16 or 32 row groups
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    bool cv[4];
    int cc[4];
    uint8_t* dst[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = (head + lane + 32 * j) & 127;
        cc[j] = c;
        cv[j] = c0 + c < ncols;
        dst[j] = wb + (c >> 6) * F_WB_CHUNK + w * 128 + ((((c & 63) >> 3) ^ w) << 4) + (c & 7) * 2;
    }
    const float* base = src + (int64_t)w * ld + c0;
    float v[32][4];
#pragma unroll
    for (int u = 0; u < 32; ++u) {
        const int i = (u + rot) & (groups - 1);
        const bool rv = u < groups && w + 8 * i < nrows;
        const float* r = base + (int64_t)(8 * i) * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[u][j] = (rv && cv[j]) ? __ldcg(r + cc[j]) : 0.f;   // L2 only: with a power-of-two pitch
                                                                                          // every row of the tile maps to the same L1 sets
    }
    mid();                                                  // with the whole tile in flight
#pragma unroll
    for (int u = 0; u < 32; ++u) {
        const int i = (u + rot) & (groups - 1);
        if (u < groups) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ss[j] = fmaf(v[u][j], v[u][j], ss[j]);
                *reinterpret_cast<__nv_bfloat16*>(dst[j] + i * 1024) = __float2bfloat16_rn(v[u][j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[w * 128 + cc[j]] = ss[j];
}

struct Smem {
    uint8_t *E, *DZ, *WB;
    float2* col;           // [128] per column: (scale * log2(e), 0 or F_NEG); scale = 1/||w_c|| (instance) or 1/T (InfoNCE), 0 if excluded
    uint64_t *bar_load, *bar_mma, *bar_dw, *bar_fin;     // bar_fin[2]: the two buffers of the final reduction
    uint32_t* tmem_slot;
    float* red32;          // [32]  block_sum scratch
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define F_STAMP(k)                                                                                     \
    do {                                                                                               \
        if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + (k)] = globaltimer_ns();     \
    } while (0)

constexpr float F_NEG = -1e30f;
constexpr float F_LOG2E = 1.4426950408889634f, F_LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {      // 2^x, MUFU.EX2; underflows to exactly 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// one full 32-byte sector per lane: no partial-sector writes (which cost a read-modify-write in L2 on ECC memory)
__device__ __forceinline__ void st_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}      // finite stand-in for -inf in the running maxima (exp underflows to exactly 0)

// base-2 log-sum-exp of one row from the per-tile values ls[tile][256] = log2 sum_c 2^(z2_c) of that tile (4 bytes per row and
// tile: this combine is bound by the bytes each SM pulls from L2, ~40 GB/s per SM); l0 = a first term (InfoNCE: the positive
// logit; F_NEG for none).  All loads of a batch are in flight together.
__device__ __forceinline__ float lse2_of_row(const float* __restrict__ ls, int row, int tiles, float l0) {
    constexpr int B = 96;
    float M = l0, S = 1.0f;                       // running max and sum of 2^(l - M); l0 = F_NEG contributes 2^(F_NEG - M) = 0 later
    for (int t0 = 0; t0 < tiles; t0 += B) {
        float a[B];
#pragma unroll
        for (int i = 0; i < B; ++i) a[i] = __ldcg(ls + (size_t)min(t0 + i, tiles - 1) * 256 + row);
        asm volatile("" ::: "memory");
        float mb = M;
#pragma unroll
        for (int i = 0; i < B; ++i) {
            a[i] = (t0 + i < tiles) ? fmaxf(a[i], F_NEG) : F_NEG;
            mb = fmaxf(mb, a[i]);
        }
        float acc = S * ex2(M - mb);
#pragma unroll
        for (int i = 0; i < B; ++i) acc += ex2(a[i] - mb);
        M = mb; S = acc;
    }
    return M + log2f(S);
}
__device__ __forceinline__ float2 sum_of_row(const float2* __restrict__ zz, int row, int tiles) {
    float sx = 0.f, sy = 0.f;
    for (int t0 = 0; t0 < tiles; t0 += 32) {
        float2 a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = (t0 + i < tiles) ? __ldcg(zz + (size_t)(t0 + i) * 256 + row) : make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; ++i) { sx += a[i].x; sy += a[i].y; }
    }
    return make_float2(sx, sy);
}

// u[i] of lane L = value of column i held by row L.  Returns, in lane L, the sum over the 32 rows of column L
// (butterfly transpose-reduce: 31 shuffles, fixed order).
__device__ __forceinline__ float lane_transpose_sum(float (&u)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const float send = up ? u[i] : u[i + step];
            const float keep = up ? u[i + step] : u[i];
            u[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return u[0];
}

// ------------------------------------------------------------------------------------------------------------------------
// prologue tasks, one warp each.  Row task: fp32 artefacts of one padded embedding row as in loss_f32.cu's prologue (same
// expressions) plus its slice of the three packed bf16 operand images (raw embeds, normalised embeds, normalised InfoNCE
// queries), zero padded.  Pack task: one full row of a fp32 [D, K] queue (contiguous, DRAM friendly; a strided 128-column tile
// read of the power-of-two-pitched queue crawls) -> bf16 into the per-tile shared-memory images
// [tile][2 column chunks][256 d][128 B] that the InfoNCE CTAs copy in one piece.
// They run as a separate small launch (fused_prologue_kernel) whose outputs reach the cooperative kernel by bulk copy while the
// W tiles stream in.  (A one-launch form -- these tasks inside the cooperative kernel -- was measured slower and removed.)
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pro_pack_task(const ProArgs& a, int mod, int d, int lane) {
    const int K = a.K, T_k = a.T_k;
    const float* row = (mod ? a.v_queue : a.t_queue) + (int64_t)d * K;       // image queries score the TEXT queue (head.py:162,168)
    uint8_t* img = a.QUp + (size_t)mod * T_k * F_WB_BYTES;
    const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0 && d < a.D;
    for (int t0 = 0; t0 < T_k; t0 += 16) {
        float4 x[16];
        if (vec) {                                 // all loads of the row in flight before the first use
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = min((t0 + i) * 128 + 4 * lane, K - 4);
                x[i] = __ldcs(reinterpret_cast<const float4*>(row + k));
            }
            asm volatile("" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if ((t0 + i) * 128 + 4 * lane + 3 >= K) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = (t0 + i) * 128 + 4 * lane;
                x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (d < a.D) {
                    if (k < K) x[i].x = row[k];
                    if (k + 1 < K) x[i].y = row[k + 1];
                    if (k + 2 < K) x[i].z = row[k + 2];
                    if (k + 3 < K) x[i].w = row[k + 3];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (t0 + i < T_k) {
                uint2 o;
                o.x = pack2(x[i].x, x[i].y); o.y = pack2(x[i].z, x[i].w);
                *reinterpret_cast<uint2*>(img + (size_t)(t0 + i) * F_WB_BYTES + wb_offset(d, 4 * lane)) = o;
            }
        }
    }
}

__device__ __forceinline__ void pro_row_task(const ProArgs& p, int task, int lane) {
    const int N = p.N, D = p.D, KC = p.KC;
    const int win = task >> 8, prow = task & 255;                  // window, padded row inside it: modality * 128 + local row
    const int mod = prow >> 7, n = (win << 7) + (prow & 127);      // n = row of the batch
    const int k0 = lane * 8;
    const bool in_pad = k0 < KC * 64;
    float a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = b[i] = c[i] = 0.f;
    const bool live = n < N && k0 < D;
    const int row = mod * N + n;
    if (live) {                                                              // 16-byte aligned rows (checked by the caller)
        const float4* e = reinterpret_cast<const float4*>((mod ? p.t_embed : p.v_embed) + (int64_t)n * D + k0);
        const float4* r = reinterpret_cast<const float4*>((mod ? p.t_qraw : p.v_qraw) + (int64_t)n * D + k0);
        const float4* kin = reinterpret_cast<const float4*>((mod ? p.v_key : p.t_key) + (int64_t)n * D + k0);   // v queries pair with TEXT keys (head.py:160,166)
        const float4 e0 = e[0], e1 = e[1], r0 = r[0], r1 = r[1], c0 = kin[0], c1 = kin[1];
        a[0] = e0.x; a[1] = e0.y; a[2] = e0.z; a[3] = e0.w; a[4] = e1.x; a[5] = e1.y; a[6] = e1.z; a[7] = e1.w;
        b[0] = r0.x; b[1] = r0.y; b[2] = r0.z; b[3] = r0.w; b[4] = r1.x; b[5] = r1.y; b[6] = r1.z; b[7] = r1.w;
        c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
    }
    float se = 0.f, sr = 0.f, sk = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { se = fmaf(a[i], a[i], se); sr = fmaf(b[i], b[i], sr); sk = fmaf(c[i], c[i], sk); }
    se = warp_sum(se); sr = warp_sum(sr); sk = warp_sum(sk);
    const float ne = fmaxf(sqrtf(se), 1e-12f), nr = fmaxf(sqrtf(sr), 1e-12f);
    const float nk = p.normalize_keys ? fmaxf(sqrtf(sk), 1e-12f) : 1.0f;
    float an[8], qv[8], kv[8], dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        an[i] = __fdiv_rn(a[i], ne);
        qv[i] = __fdiv_rn(b[i], nr);
        kv[i] = p.normalize_keys ? __fdiv_rn(c[i], nk) : c[i];
        dot = fmaf(qv[i], kv[i], dot);
    }
    dot = warp_sum(dot);
    if (live) {
        float4* kout = reinterpret_cast<float4*>((mod ? p.v_key_n : p.t_key_n) + (int64_t)n * D + k0);
        float4* o_e = reinterpret_cast<float4*>(p.E2 + (int64_t)row * D + k0);
        float4* o_en = reinterpret_cast<float4*>(p.en + (int64_t)row * D + k0);
        float4* o_qn = reinterpret_cast<float4*>(p.qn + (int64_t)row * D + k0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            o_e[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
            o_en[i] = make_float4(an[4 * i], an[4 * i + 1], an[4 * i + 2], an[4 * i + 3]);
            o_qn[i] = make_float4(qv[4 * i], qv[4 * i + 1], qv[4 * i + 2], qv[4 * i + 3]);
            kout[i] = make_float4(kv[4 * i], kv[4 * i + 1], kv[4 * i + 2], kv[4 * i + 3]);
        }
    }
    if (n < N && lane == 0) { p.inv_e[row] = __fdiv_rn(1.0f, ne); p.inv_q[row] = __fdiv_rn(1.0f, nr); p.pos[row] = dot; }
    if (in_pad) {
        const int64_t off = (int64_t)win * 2 * KC * BLOCK_BYTES + packed_offset_bytes(prow, lane, KC);
        uint4 o;
        o.x = pack2(a[0], a[1]); o.y = pack2(a[2], a[3]); o.z = pack2(a[4], a[5]); o.w = pack2(a[6], a[7]);
        *reinterpret_cast<uint4*>(p.Ep + off) = o;
        o.x = pack2(an[0], an[1]); o.y = pack2(an[2], an[3]); o.z = pack2(an[4], an[5]); o.w = pack2(an[6], an[7]);
        *reinterpret_cast<uint4*>(p.ENp + off) = o;
        o.x = pack2(qv[0], qv[1]); o.y = pack2(qv[2], qv[3]); o.z = pack2(qv[4], qv[5]); o.w = pack2(qv[6], qv[7]);
        *reinterpret_cast<uint4*>(p.QNp + off) = o;
    }
}

// task t of the prologue: 256 padded embedding rows per window, then the queue rows (2 * Dp) when `pack`
__device__ __forceinline__ void pro_task(const ProArgs& a, int t, int lane) {
    const int nrow = 256 * ((a.N + 127) / 128);
    if (t < nrow) pro_row_task(a, t, lane);
    else { const int q = t - nrow, Dp = a.KC * 64; pro_pack_task(a, q / Dp, q % Dp, lane); }
}

// stand-alone prologue (some branches unfused): one warp per task, also clears the grid-barrier words
__global__ void __launch_bounds__(256) fused_prologue_kernel(const ProArgs a, unsigned* __restrict__ bar, int ntasks) {
    griddep_launch_dependents();         // the cooperative kernel may start streaming W now; it waits for this grid before it reads
    // W towards L2 as ONE sequential sweep (bulk prefetch, a slice per CTA): the instance tiles read it as 256 row segments of
    // 512 bytes each, a pattern that pulls from cold DRAM at a quarter of the rate of a contiguous stream
    if (a.W != nullptr && threadIdx.x == 0) {
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(a.W) + 15) & ~uintptr_t(15);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(a.W) + (uintptr_t)a.W_bytes) & ~uintptr_t(15);
        if (hi > lo) {
            const uintptr_t per = (((hi - lo) / gridDim.x) + 16) & ~uintptr_t(15);
            const uintptr_t b0 = lo + per * blockIdx.x;
            if (b0 < hi) {
                const uint32_t n = (uint32_t)((hi - b0 < per) ? hi - b0 : per);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b0), "r"(n) : "memory");
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) bar[threadIdx.x] = 0u;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t < ntasks) pro_task(a, t, threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------------------------------------------
// instance (INST) / InfoNCE tile
// ------------------------------------------------------------------------------------------------------------------------
// `mma_phase` / `load_phase`: parities of the CTA's MMA and operand-load mbarriers, carried across calls (an InfoNCE CTA in dual
// mode runs the program twice, once per modality)
template <bool INST>
__device__ __forceinline__ void tile_program(const FP& p, const Smem& sm, uint32_t tmem, int mod, int tile, uint32_t& mma_phase,
                                             uint32_t& load_phase) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, q = w & 3, h = w >> 2;
    const int n = q * 32 + lane;                       // row inside a 128-row block = TMEM lane
    const int Dp = p.KC * 64;
    const int c0 = tile * F_TILE;
    constexpr int MT = INST ? 2 : 1;
    const int NW = (p.N + 127) / 128;                  // 128-row windows of the batch, processed one after the other
    const size_t img_bytes = (size_t)2 * p.KC * BLOCK_BYTES;            // packed operand image of one window
    const int ncols = INST ? p.C : p.K;
    const int tiles = INST ? p.T_inst : p.T_k;
    const uint32_t lanes = (uint32_t)(q * 32) << 16;
    // scratch inside the logit-gradient region while no gradient is staged there
    float* red = reinterpret_cast<float*>(sm.DZ);                        // [8][128] column partials
    int64_t* s_lab = reinterpret_cast<int64_t*>(sm.DZ + 8192);           // [<= 1024] ids of the whole batch (InfoNCE mask)
    float* s_lse2 = reinterpret_cast<float*>(sm.DZ + 16384);             // [256] base-2 row log-sum-exp

    // operand rows (and, InfoNCE, the queue tile image) are written by the prologue launch, which may still be running
    // (programmatic dependent launch): wait for it, then thread 0 starts the bulk copies (TMA engine).  The instance tiles
    // do this with their whole W tile already in flight.
    // operand rows of window `win` -> shared memory (thread 0; the first window also brings the InfoNCE queue tile)
    auto load_rows = [&](int win) {
        if (tid == 0) {
            const uint8_t* src = (INST ? p.Ep : p.QNp + (size_t)mod * p.KC * BLOCK_BYTES) + (size_t)win * img_bytes;
            const int blocks = MT * p.KC;
            mbar_expect_tx(sm.bar_load, (uint32_t)blocks * BLOCK_BYTES + ((INST || win > 0) ? 0u : (uint32_t)F_WB_BYTES));
#pragma unroll 1
            for (int b = 0; b < blocks; ++b) bulk_g2s(sm.E + (size_t)b * BLOCK_BYTES, src + (size_t)b * BLOCK_BYTES, BLOCK_BYTES, sm.bar_load);
            if (!INST && win == 0) {
                const uint8_t* qsrc = p.QUp + ((size_t)mod * p.T_k + tile) * F_WB_BYTES;
#pragma unroll 1
                for (int b = 0; b < F_WB_BYTES / BLOCK_BYTES; ++b)
                    bulk_g2s(sm.WB + (size_t)b * BLOCK_BYTES, qsrc + (size_t)b * BLOCK_BYTES, BLOCK_BYTES, sm.bar_load);
            }
        }
    };
    auto after_prologue = [&]() {
        griddep_wait();
        load_rows(0);
    };
    int64_t slot_id = -1;
    if (!INST) {
        for (int i = tid; i < ((p.N + 7) & ~7); i += F_THREADS)               // up to 1024 batch ids (8 KB of the region)
            s_lab[i] = i < p.N ? p.labels[i] : INT64_MIN;
        if (tid < 128 && c0 + tid < ncols) slot_id = p.id_queue[c0 + tid];
        after_prologue();
    }

    // ---- instance: W tile HBM -> bf16 shared image, column norms
    if (INST) load_tile_bf16(p.W, ncols, p.D, Dp, c0, ncols, (tile * 7) & 31, sm.WB, red, after_prologue);
    __syncthreads();
    if (tid < 128) {
        bool ok = c0 + tid < ncols;
        float scale = 0.f;
        if (INST) {
            float tot = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) tot += red[ww * 128 + tid];
            if (ok) scale = __fdiv_rn(1.0f, fmaxf(sqrtf(tot), 1e-12f));         // losses.py:51
        } else {
            if (ok) {                                                           // head.py:148-157: drop slots holding a batch id
                bool hit = false;
                const int nm = (p.N + 7) & ~7;
#pragma unroll 8
                for (int i = 0; i < nm; ++i) hit |= (s_lab[i] == slot_id);
                ok = !hit;
            }
            if (ok) scale = __fdiv_rn(1.0f, p.T);
        }
        sm.col[tid] = make_float2(scale * F_LOG2E, ok ? 0.f : F_NEG);
    }
    fence_async_smem();
    __syncthreads();
    F_STAMP(1);

    const bool do_dw = INST && p.want_grad && p.d_proj != nullptr;
    const int NH = Dp / 128;
    float csum0 = 0.f, csum1 = 0.f;                    // <dz', z2>_rows of this thread's two 32-column chunks (column = lane)
    // ---- the batch in windows of 128 rows per modality: the W / queue tile, its column scales and the dWs accumulator (TMEM)
    //      stay resident, only the operand rows change.  One window for the BASELINE batch of 128.
#pragma unroll 1
  for (int win = 0; win < NW; ++win) {
    const int r0 = win * 128;
    const int N = min(128, p.N - r0);                  // rows of this window
    const size_t stat_off = (size_t)win * tiles * 256;
    if (win > 0) load_rows(win);                       // every MMA that read the previous rows has completed (waited below)

    // ---- forward logits: z[mt] = E[mt] . Wb   (M = 128 rows, N = 128 columns, K = Dp)
    if (tid == 0) {
        mbar_wait(sm.bar_load, load_phase);
        tc_fence_after();
        const uint32_t id_f = idesc(128, 128, 0, 1);
        const uint32_t e0 = smem_u32(sm.E), wb0 = smem_u32(sm.WB);
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll 1
            for (int ks = 0; ks < Dp / 16; ++ks)
                umma_bf16(tmem + mt * 128, umma_desc_sw128(e0 + (mt * p.KC + (ks >> 2)) * BLOCK_BYTES + (ks & 3) * 32),
                          desc_mn(wb0 + ks * 2048, F_WB_CHUNK), id_f, (uint32_t)(ks > 0));
        umma_commit(sm.bar_mma);
    }
    mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
    mma_phase ^= 1;
    load_phase ^= 1;                                   // one operand load per window
    tc_fence_after();
    F_STAMP(2);

    // ---- debug: the logits of one instance tile leave the SM (never on the product path: dbg_logits is NULL)
    if (INST && p.dbg_logits != nullptr && tile == p.dbg_tile && h < MT && win == 0) {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float v[32];
            tmem_ld32(tmem + lanes + (uint32_t)(h * 128 + j * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) p.dbg_logits[(size_t)(h * 128 + n) * 128 + j * 32 + i] = v[i] * (sm.col[j * 32 + i].x * F_LN2);
        }
    }
    // ---- per-row partial softmax statistics of this tile, base 2 (thread = row n of block h): z2 = acc * scale * log2(e)
    // a label outside [0, C) makes the reference raise (scatter_ / CrossEntropyLoss); here the row's loss becomes NaN
    const int64_t lab_n = INST ? p.labels[r0 + (n < N ? n : 0)] : 0;
    const bool bad_label = INST && (lab_n < 0 || lab_n >= (int64_t)p.C);
    const int y = INST ? (bad_label ? -(1 << 30) : (int)lab_n - c0) : -1;
    if (h < MT) {
        float m = F_NEG, s = 0.f, sz2 = 0.f, zy2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float v[32];
            tmem_ld32(tmem + lanes + (uint32_t)(h * 128 + j * 32), v);
            float cm = F_NEG;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float2 cc = sm.col[j * 32 + i];
                sz2 = fmaf(v[i], cc.x, sz2);                       // excluded columns have scale 0
                v[i] = fmaf(v[i], cc.x, cc.y);                     // ... and bias F_NEG
                cm = fmaxf(cm, v[i]);
            }
            if (INST && (unsigned)(y - j * 32) < 32u) {            // rare: this row's label lies in this chunk
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i == y - j * 32) zy2 = v[i];
            }
            const float nm = fmaxf(m, cm);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += ex2(v[i] - nm);
            s = s * ex2(m - nm) + acc;
            m = nm;
        }
        if (n < N) {
            const size_t at = stat_off + (size_t)tile * 256 + (INST ? h : mod) * 128 + n;
            (INST ? p.ls_inst : p.ls_nce)[at] = m + log2f(s);      // s >= 1 (the maximum contributes 2^0)
            if (INST) p.zz_inst[at] = make_float2(sz2 * F_LN2, zy2 * F_LN2);
        }
    }
    F_STAMP(3);

    // the row statistics of a loss only involve its own tiles: one counter per group (instance; InfoNCE per modality), so the
    // InfoNCE tiles never wait for the (later) instance tiles
    {
        unsigned* ctr = INST ? p.bar : p.bar + 2 + mod;
        grid_arrive(ctr);
        grid_wait(ctr, (unsigned)(tiles * (win + 1)));      // one arrival per tile and window
    }
    F_STAMP(4);

    // ---- row log-sum-exp over all tiles (one thread per row, shared through shared memory); the first tile writes the row losses
    if (h < MT && n < N) {
        if (INST) {
            const int row = h * 128 + n;
            const float l2 = lse2_of_row(p.ls_inst + stat_off, row, tiles, F_NEG);
            s_lse2[row] = l2;
            if (tile == 0 && p.row_helpers == 0) {                              // losses.py:26-39 with label smoothing (else: spare CTAs)
                const float2 zz = sum_of_row(p.zz_inst + stat_off, row, tiles);
                p.rows_inst[h * p.NS + r0 + n] = bad_label ? CUDART_NAN_F : l2 * F_LN2 - (1.0f - p.eps) * zz.y - (p.eps / (float)p.C) * zz.x;
            }
        } else {
            const float z02 = __fdiv_rn(p.pos[mod * p.NS + r0 + n], p.T) * F_LOG2E;     // column 0 of the reference's logits (base 2)
            const float l2 = lse2_of_row(p.ls_nce + stat_off, mod * 128 + n, tiles, z02);
            s_lse2[n] = l2;
            if (tile == 0) {                                                    // losses.py:206-217, target 0
                p.rows_nce[mod * p.NS + r0 + n] = (l2 - z02) * F_LN2;                // exactly 0 when only the positive is left
                p.dpos[mod * p.NS + r0 + n] = (exp2f(z02 - l2) - 1.0f) / ((float)p.Nn * p.T);
            }
        }
    }
    __syncthreads();
    float lse2[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) lse2[mt] = n < N ? s_lse2[mt * 128 + n] : 0.f;
    __syncthreads();                                    // the region is rewritten with logit gradients below
    F_STAMP(5);

    if (p.want_grad) {
        const float rs = n < N ? F_LN2 / (float)p.Nn : 0.f;          // dz' = (softmax - target) / N * scale, scale = col.x * ln 2
        const float uni = INST ? p.eps / (float)p.C : 0.f;
        const float uni_hot = uni + (INST ? 1.0f - p.eps : 0.f);
        // TMEM columns: Z0 = [0,128) and Z1 = [128,256) hold the logits of row blocks 0 / 1, W0|W1 = [256,512) the dWs accumulator.
        // Row block 0: dE halves go to Z0 (its logits are consumed) and W0 (dWs has not started); once they are drained, dWs(0)
        // is issued and runs under the gradient math of row block 1.  Row block 1: dE halves go to Z1 and Z0, dWs(1)
        // accumulates on top; one wait, one drain.
        // Real loops (no unrolling over row blocks / chunks): the kernel is instruction-fetch bound when this is straight-line
        // code (ncu: `no instruction` was the top stall reason), so every body below is fetched once and re-used.
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const float lse_mt = mt == 0 ? lse2[0] : lse2[MT - 1];
            // -- logit gradient of row block mt -> bf16 (thread = row n, 64 columns of chunk h)
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int j = h * 2 + jj;
                const int yy = y - j * 32;
                float v[32], u[32];
                tmem_ld32(tmem + lanes + (uint32_t)(mt * 128 + j * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float sc2 = sm.col[j * 32 + i].x;
                    const float z2 = v[i] * sc2;
                    const float g = ex2(z2 - lse_mt) - ((i == yy) ? uni_hot : uni);
                    v[i] = g * (sc2 * rs);                         // exactly 0 for excluded columns and padding rows
                    u[i] = v[i] * z2;
                }
                uint4 o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[k].x = pack2(v[8 * k + 0], v[8 * k + 1]); o[k].y = pack2(v[8 * k + 2], v[8 * k + 3]);
                    o[k].z = pack2(v[8 * k + 4], v[8 * k + 5]); o[k].w = pack2(v[8 * k + 6], v[8 * k + 7]);
                }
                // <dWs, What>_col = sum_rows dz'[row, c] * z[row, c]  (dWs = E^T dz', z = E What): the column-normalisation
                // Jacobian needs no second pass over W
                if (do_dw) {
                    const float t = lane_transpose_sum(u, lane);
                    if (jj) csum1 += t; else csum0 += t;
                }
                if (mt == 1 && jj == 0 && do_dw) {     // dWs(0) still reads the previous image
                    mbar_wait_sleepy(sm.bar_dw, (uint32_t)(win & 1), 32);
                    tc_fence_after();
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(sm.DZ + dz_offset(h, n, jj * 4 + k)) = o[k];
            }
            tc_fence_before();
            fence_async_smem();
            __syncthreads();
            F_STAMP(10 + 2 * mt);

            const bool two = INST && NH == 2;          // dE has two 128-column halves
            const uint32_t colA = INST ? (uint32_t)(mt * 128) : 256u;            // half a (InfoNCE: all Dp columns)
            // half b: row block 0 borrows W0 while dWs has not started -- from the second window on W0 holds the accumulated
            // dWs, and the two halves of row block 0 go through Z0 one after the other
            const bool serial = two && mt == 0 && win > 0;
            const uint32_t colB = serial ? colA : (mt == 0 ? 256u : 0u);
            const int halves = two ? 2 : 1;
#pragma unroll 1
            for (int pass = 0; pass < (serial ? 2 : 1); ++pass) {
                const int r_lo = serial ? pass : 0, r_hi = serial ? pass + 1 : halves;
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t e0 = smem_u32(sm.E), wb0 = smem_u32(sm.WB), dz0 = smem_u32(sm.DZ);
                    // dE[row, d] = sum_c dz'[row, c] * Wb[d, c]           (M = 128 rows, N = d, K = 128 columns)
                    const uint32_t id_e = idesc(128, INST ? 128 : Dp, 0, 0);
#pragma unroll 1
                    for (int r = r_lo; r < r_hi; ++r)
#pragma unroll 1
                        for (int ks = 0; ks < 8; ++ks)
                            umma_bf16(tmem + (r ? colB : colA), umma_desc_sw128(dz0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32),
                                      umma_desc_sw128(wb0 + (ks >> 2) * F_WB_CHUNK + r * (128 * 128) + (ks & 3) * 32), id_e,
                                      (uint32_t)(ks > 0));
                    if (do_dw && mt == 1) {
                        // dWs[d, c] += sum_rows E[row, d] * dz'[row, c]   (M = 128 d per half, N = 128 columns, K = 128 rows)
                        const uint32_t id_w = idesc(128, 128, 1, 1);
                        for (int hh = 0; hh < NH; ++hh)
#pragma unroll 1
                            for (int ks = 0; ks < 8; ++ks)
                                umma_bf16(tmem + 256 + hh * 128,
                                          desc_mn(e0 + (p.KC + 2 * hh) * BLOCK_BYTES + ks * 2048, BLOCK_BYTES),
                                          desc_mn(dz0 + ks * 2048, BLOCK_BYTES), id_w, 1u);
                    }
                    umma_commit(sm.bar_mma);
                }
                mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
                mma_phase ^= 1;
                tc_fence_after();
                // -- drain the partial dE of this tile to global (reduced over the tiles at the end of the kernel); partials are
                // rounded to bf16 (their own error from the bf16 operands is 2^-8; they are summed in fp32)
                if (INST) {
                    // slot m = (row block * Dp/8 + 8-column chunk) * 128 + row; layout [window][unit = m / U][tile][m % U] (U
                    // divides 128): the share of all tiles in one work unit of the final reduction is ONE contiguous piece (see
                    // PartialReducer); the 32 rows of a warp write 512 contiguous bytes
                    const int U = p.fin_U, upr = 128 / U;                    // units per 128-row group
                    const size_t win_off = (size_t)win * p.T_inst * (2 * (Dp / 8) * 128);
                    uint4* dst = p.part_inst + win_off + ((size_t)(n / U) * p.T_inst + tile) * U + (n % U);
                    const size_t gstride = (size_t)upr * p.T_inst * U;       // one (row block, chunk) group further
#pragma unroll 1
                    for (int rj = 2 * r_lo; rj < 2 * r_hi; ++rj) {
                        const int r = rj >> 1, jj = rj & 1;
                        float v[32];
                        tmem_ld32(tmem + lanes + (r ? colB : colA) + (uint32_t)(h * 64 + jj * 32), v);
                        if (n < N) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                __stcg(dst + (size_t)(mt * (Dp / 8) + r * 16 + h * 8 + jj * 4 + k) * gstride,
                                       make_uint4(pack2(v[8 * k], v[8 * k + 1]), pack2(v[8 * k + 2], v[8 * k + 3]),
                                                  pack2(v[8 * k + 4], v[8 * k + 5]), pack2(v[8 * k + 6], v[8 * k + 7])));
                        }
                    }
                } else {
                    const int half = Dp / 2;
                    // [window][modality][tile][row][8-column chunk]: the reader (one warp per row, lane = chunk) pulls 512
                    // contiguous bytes per tile; a thread writes 64 contiguous bytes per TMEM chunk
                    uint4* dst = p.part_nce + ((size_t)win * 2 * p.T_k + (size_t)(mod * p.T_k + tile)) * (Dp / 8) * 128 + (size_t)n * (Dp / 8);
#pragma unroll 1
                    for (int jj = 0; jj < half / 32; ++jj) {
                        float v[32];
                        tmem_ld32(tmem + lanes + (uint32_t)(256 + h * half + jj * 32), v);
                        if (n < N) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                __stcg(dst + ((h * half + jj * 32) / 8 + k),
                                       make_uint4(pack2(v[8 * k], v[8 * k + 1]), pack2(v[8 * k + 2], v[8 * k + 3]),
                                                  pack2(v[8 * k + 4], v[8 * k + 5]), pack2(v[8 * k + 6], v[8 * k + 7])));
                        }
                    }
                }
                if (serial && pass == 0) {             // Z0 is re-used by the second half
                    tc_fence_before();
                    __syncthreads();
                }
            }
            tc_fence_before();
            __syncthreads();                           // the drained TMEM columns may be overwritten
            F_STAMP(11 + 2 * mt);
            if (do_dw && mt == 0 && tid == 0) {
                tc_fence_after();
                const uint32_t e0 = smem_u32(sm.E), dz0 = smem_u32(sm.DZ);
                const uint32_t id_w = idesc(128, 128, 1, 1);
                for (int hh = 0; hh < NH; ++hh)
#pragma unroll 1
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem + 256 + hh * 128, desc_mn(e0 + (2 * hh) * BLOCK_BYTES + ks * 2048, BLOCK_BYTES),
                                  desc_mn(dz0 + ks * 2048, BLOCK_BYTES), id_w, (uint32_t)(win > 0 || ks > 0));
                umma_commit(sm.bar_dw);
            }
        }
    }      // want_grad
  }        // windows
    if (p.want_grad) {
        F_STAMP(6);
        // the partial dE tiles are out: the CTAs that reduce them (InfoNCE / align / spare CTAs, see fused_loss_kernel) can start
        // while this CTA still writes its dW tile
        if (INST && p.fin_early) grid_arrive(p.bar + 1);

        if (do_dw) {
            // ---- dW tile = dWs - What * <dWs, What>_col   (dWs already carries the 1/||w|| factor)       losses.py:51
            // TMEM [d lanes, 128 columns] -> fp32 shared tile [Dp][128] (float4 index XOR (d & 7)) in the embedding region
            float4* tile4 = reinterpret_cast<float4*>(sm.E);
            const float* tile1 = reinterpret_cast<const float*>(sm.E);
            if (h < NH) {
                const int d = h * 128 + n;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    float v[32];
                    tmem_ld32(tmem + lanes + (uint32_t)(256 + h * 128 + j * 32), v);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        tile4[d * 32 + ((j * 8 + k) ^ (d & 7))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                }
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) red[q * 128 + (h * 2 + jj) * 32 + lane] = jj ? csum1 : csum0;
            __syncthreads();
            F_STAMP(14);
            bool cv[4];
            int cc[4];
            float sc[4];
            const int head = sector_head(w, p.C, c0);  // full, aligned 128-byte segments: no partial-sector writes inside the tile
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = (head + lane + 32 * j) & 127;
                cc[j] = c;
                cv[j] = c0 + c < p.C;
                // What * <dWs, What> = W * (scale * dot); csum is in base-2 units: dot = csum * ln 2, scale = col.x * ln 2
                sc[j] = (sm.col[c].x * F_LN2) * ((((red[c] + red[128 + c]) + red[256 + c]) + red[384 + c]) * F_LN2);
            }
            // What comes from the bf16 shared image of the tile (relative error 2^-9 on the subtracted projection only)
            const uint8_t* wsrc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = cc[j];
                wsrc[j] = sm.WB + (c >> 6) * F_WB_CHUNK + w * 128 + ((((c & 63) >> 3) ^ w) << 4) + (c & 7) * 2;
            }
            for (int i0 = 0; i0 < p.D / 8; i0 += 8) {   // 8 rows x 4 columns per batch: shared-memory reads first, then 32 stores
                float o[8][4];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u, d = w + 8 * i;   // (d & 7) == w, (d >> 3) == i
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = cc[j];
                        const float wv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(wsrc[j] + i * 1024));
                        o[u][j] = fmaf(-wv, sc[j], tile1[d * 128 + ((((c >> 2) ^ w)) << 2) + (c & 3)]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int d = w + 8 * (i0 + u);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (cv[j]) p.d_proj[(int64_t)d * p.C + c0 + cc[j]] = o[u][j];
                }
            }
        }
        F_STAMP(7);
    } else if (INST && p.fin_early) {
        grid_arrive(p.bar + 1);
    }
}

// 8 consecutive bf16 (one 16-byte chunk) of embedding row `row`, elements d0 .. d0+7, from a packed row block in shared memory
__device__ __forceinline__ void smem_row_chunk(const uint8_t* blk0, int row, int d0, float (&out)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(blk0 + (d0 >> 6) * BLOCK_BYTES + (row >> 3) * 1024 + (row & 7) * 128 +
                                                      ((((d0 & 63) >> 3) ^ (row & 7)) << 4));
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        out[2 * k] = __uint_as_float(r[k] << 16);
        out[2 * k + 1] = __uint_as_float(r[k] & 0xffff0000u);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// global align (losses.py:102-128): S = en_v en_t^T, pair losses, dS, dq_v = dS en_t, dq_t = dS^T en_v, normalise backward.
// One CTA; the bf16-rounded normalised embeddings in shared memory are used consistently (MMA operands and projection).
// ------------------------------------------------------------------------------------------------------------------------
// Batches above 128 rows: one align CTA per 128-row window `iw` of the IMAGE rows; it walks the 128-row windows `jw` of the text
// rows (block S[iw, jw] at a time), accumulates dq_v[iw] in TMEM over them and hands the dq_t contribution of every block to
// global memory; the reducing CTAs sum them and finish the text rows (finish_phase, one warp per row).
__device__ __forceinline__ void align_program(const FP& p, const Smem& sm, uint32_t tmem, int iw) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, q = w & 3, h = w >> 2;
    const int n = q * 32 + lane;
    const int Nall = p.N, D = p.D, Dp = p.KC * 64;
    const int NW = (Nall + 127) / 128;
    const int ri = iw * 128, Ni = min(128, Nall - ri);                  // image rows of this CTA
    const uint32_t lanes = (uint32_t)(q * 32) << 16;
    const size_t img_bytes = (size_t)2 * p.KC * BLOCK_BYTES;           // packed image of one window: [v rows][t rows]
    uint32_t mma_phase = 0, load_phase = 0;
    int64_t* s_labi = reinterpret_cast<int64_t*>(sm.WB);               // [128] ids of the image rows
    int64_t* s_labj = reinterpret_cast<int64_t*>(sm.WB + 1024);        // [128] ids of the text rows of the current block
    float* s_part = reinterpret_cast<float*>(sm.WB + 2048);            // [2][128]
    const uint32_t e0 = smem_u32(sm.E), dz0 = smem_u32(sm.DZ);
    const uint32_t et0 = e0 + p.KC * BLOCK_BYTES;                      // text rows
    const int half = Dp / 2;

    griddep_wait();
    if (tid < 128) s_labi[tid] = tid < Ni ? p.labels[ri + tid] : INT64_MIN;
    float acc = 0.f;                                                   // pair losses of (row n, column half h) over all blocks
#pragma unroll 1
    for (int jw = 0; jw < NW; ++jw) {
        const int rj = jw * 128, Nj = min(128, Nall - rj);
        if (tid == 0) {                                                // (every MMA that read the previous text rows has completed)
            const int first = jw == 0 ? 0 : p.KC;                      // the image rows arrive once
            mbar_expect_tx(sm.bar_load, (uint32_t)(2 * p.KC - first) * BLOCK_BYTES);
#pragma unroll 1
            for (int b = first; b < 2 * p.KC; ++b) {
                const int win = b < p.KC ? iw : jw;
                bulk_g2s(sm.E + (size_t)b * BLOCK_BYTES, p.ENp + (size_t)win * img_bytes + (size_t)b * BLOCK_BYTES, BLOCK_BYTES, sm.bar_load);
            }
        }
        if (tid < 128) s_labj[tid] = tid < Nj ? p.labels[rj + tid] : INT64_MIN;
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            mbar_wait(sm.bar_load, load_phase);
            tc_fence_after();
            const uint32_t id_s = idesc(128, 128, 0, 0);
#pragma unroll 1
            for (int ks = 0; ks < Dp / 16; ++ks)
                umma_bf16(tmem, umma_desc_sw128(e0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32),
                          umma_desc_sw128(et0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32), id_s, (uint32_t)(ks > 0));
            umma_commit(sm.bar_mma);
        }
        mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
        mma_phase ^= 1;
        load_phase ^= 1;
        tc_fence_after();
        F_STAMP(2);

        {
            const int64_t yi = s_labi[n];
            const float two_over_n = 2.0f / (float)Nall;
            const float c_same = -p.sp * two_over_n, c_diff = p.sn * two_over_n;
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int j = h * 2 + jj;
                float v[32];
                tmem_ld32(tmem + lanes + (uint32_t)(j * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int col = j * 32 + i;
                    const bool ok = n < Ni && col < Nj;
                    const bool same = s_labj[col] == yi;
                    const float x = same ? -p.sp * (v[i] - p.alpha) : p.sn * (v[i] - p.beta);
                    const float e = __expf(x);
                    const float ope = 1.0f + e;
                    acc += ok ? __logf(ope) : 0.f;                       // the reference's literal log(1+exp(x)), losses.py:123-124
                    v[i] = ok ? (same ? c_same : c_diff) * __fdividef(e, ope) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint4 o;
                    o.x = pack2(v[8 * k + 0], v[8 * k + 1]); o.y = pack2(v[8 * k + 2], v[8 * k + 3]);
                    o.z = pack2(v[8 * k + 4], v[8 * k + 5]); o.w = pack2(v[8 * k + 6], v[8 * k + 7]);
                    *reinterpret_cast<uint4*>(sm.DZ + dz_offset(h, n, jj * 4 + k)) = o;
                }
            }
        }
        tc_fence_before();
        fence_async_smem();
        __syncthreads();
        if (!p.want_grad) continue;

        if (tid == 0) {
            tc_fence_after();
            // dq_v[i, d] += sum_j dS[i, j] en_t[j, d]   -> columns [256, 256 + Dp), accumulated over the text windows
            const uint32_t id_v = idesc(128, Dp, 0, 1);
#pragma unroll 1
            for (int ks = 0; ks < 8; ++ks)
                umma_bf16(tmem + 256, umma_desc_sw128(dz0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32),
                          desc_mn(et0 + ks * 2048, BLOCK_BYTES), id_v, (uint32_t)(jw > 0 || ks > 0));
            // dq_t[j, d] = sum_i dS[i, j] en_v[i, d]   -> columns [0, Dp): this block's contribution
            const uint32_t id_t = idesc(128, Dp, 1, 1);
#pragma unroll 1
            for (int ks = 0; ks < 8; ++ks)
                umma_bf16(tmem, desc_mn(dz0 + ks * 2048, BLOCK_BYTES), desc_mn(e0 + ks * 2048, BLOCK_BYTES), id_t,
                          (uint32_t)(ks > 0));
            umma_commit(sm.bar_mma);
        }
        mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
        mma_phase ^= 1;
        tc_fence_after();
        if (NW > 1) {
            // several image windows contribute to these text rows: fp32 block [128 rows][Dp] to global, summed after the counter
            float* part = p.ga_part + ((size_t)(iw * NW + jw) * 128 + n) * Dp + h * half;
#pragma unroll 1
            for (int jj = 0; jj < half / 32; ++jj) {
                float v[32];
                tmem_ld32(tmem + lanes + (uint32_t)(h * half + jj * 32), v);
#pragma unroll
                for (int k = 0; k < 4; ++k) st_v8(part + jj * 32 + 8 * k, &v[8 * k]);
            }
            tc_fence_before();
            __syncthreads();                                             // S of the next block overwrites these columns
        }
    }
    s_part[h * 128 + n] = acc;
    __syncthreads();
    if (h == 0 && n < Ni) p.rows_ga[ri + n] = s_part[n] + s_part[128 + n];
    __syncthreads();
    F_STAMP(3);
    // nothing the reducing CTAs need comes after the row losses: arrive now, the rest runs beside the reductions
    if (p.fin_early) grid_arrive(p.bar + 1);

    if (p.want_grad) {
        F_STAMP(5);
        // normalise backward: d_ga[row] = (g - <g, en> en) / ||e||, thread = (row n, column half h)
        // image rows of window iw: g = dq_v in TMEM; text rows of a single-window batch: dq_t in TMEM
        // (with several windows the text rows are finished by the reducing CTAs, one warp per row: finish_phase)
        for (int mod = 0; mod < (NW == 1 ? 2 : 1); ++mod) {
            const uint32_t col0 = (mod == 0 ? 256u : 0u) + (uint32_t)(h * half);
            const uint8_t* blk0 = sm.E + (size_t)mod * p.KC * BLOCK_BYTES;
            auto load_g = [&](int jj, float (&v)[32]) { tmem_ld32(tmem + lanes + col0 + (uint32_t)(jj * 32), v); };
            const int Nm = Ni;                                           // rows of window iw in either modality
            float dot = 0.f;
#pragma unroll 1
            for (int jj = 0; jj < half / 32; ++jj) {
                float v[32];
                load_g(jj, v);
                const int d0 = h * half + jj * 32;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float e8[8];
                    smem_row_chunk(blk0, n, d0 + 8 * k, e8);
#pragma unroll
                    for (int t = 0; t < 8; ++t) dot = fmaf(v[8 * k + t], e8[t], dot);
                }
            }
            __syncthreads();
            s_part[h * 128 + n] = dot;
            __syncthreads();
            dot = s_part[n] + s_part[128 + n];
            const float inv = n < Nm ? p.inv_e[mod * Nall + ri + n] : 0.f;
            float* out = p.d_ga + ((int64_t)mod * Nall + ri + (n < Nm ? n : 0)) * D;
#pragma unroll 1
            for (int jj = 0; jj < half / 32; ++jj) {
                float v[32];
                load_g(jj, v);
                const int d0 = h * half + jj * 32;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float e8[8], o[8];
                    smem_row_chunk(blk0, n, d0 + 8 * k, e8);
#pragma unroll
                    for (int t = 0; t < 8; ++t) o[t] = (v[8 * k + t] - dot * e8[t]) * inv;
                    if (n < Nm && d0 + 8 * k < D) st_v8(out + d0 + 8 * k, o);
                }
            }
        }
        F_STAMP(6);
    }
}

// d_inst[row, 8 columns] = sum over the instance tiles of their partial dE tiles, in a fixed order.  One tile's partial is
// M = 2 * (Dp/8) * 128 16-byte slots m = (row block * Dp/8 + 8-column chunk) * 128 + row.  Work unit = U consecutive slots; the
// tiles store their partials as [unit][tile][U slots], so a unit's share of EVERY tile is one contiguous piece of T * U * 16
// bytes that a single bulk copy (TMA engine) lands in shared memory; FIN_PARTS threads per slot then add their tiles and the
// partial sums are combined through shared memory.  Units are handed out by a grid-wide counter (bar[4]) and double buffered:
// the phase is bound by the bytes an SM can pull from L2 (~40 GB/s each), so every CTA that gets here -- early or late -- keeps
// taking units until none is left, and the copy of the next unit runs under the additions of the current one.  Which CTA
// reduces a unit does not change the order of any sum: results are bit-identical from call to call.
struct PartialReducer {
    const FP& p;
    const Smem& sm;
    int tid, Dp, M, U, upw, n_units, cur, nxt, k;
    uint32_t ph0, ph1;              // mbarrier phases of the two buffers (scalars: no dynamically indexed local array)
    int* s_next;
    __device__ __forceinline__ PartialReducer(const FP& p_, const Smem& sm_) : p(p_), sm(sm_) {
        tid = threadIdx.x;
        Dp = p.KC * 64;
        M = 2 * (Dp / 8) * 128;
        U = p.fin_U;
        upw = M / U;                                  // units per 128-row window
        n_units = upw * ((p.N + 127) / 128);
        s_next = reinterpret_cast<int*>(sm.red32 + 32);
        cur = nxt = n_units; k = 0;
        ph0 = ph1 = 0;
    }
    __device__ __forceinline__ uint8_t* buffer(int which) const { return sm.E + (size_t)which * (F_OFF_MISC / 2); }
    // all threads: take the next unit (thread 0 asks the counter, the answer goes round through shared memory) and start its
    // copy into buffer `which`; the barriers also order every earlier generic access of that buffer before the async-proxy writes
    __device__ __forceinline__ int grab_and_copy(int which) {
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            const int unit = (int)atomicAdd(p.bar + 4, 1u);
            *s_next = unit;
            if (unit < n_units) {
                const uint32_t bytes = (uint32_t)(p.T_inst * U * 16);
                mbar_expect_tx(sm.bar_fin + which, bytes);
                bulk_g2s(buffer(which), p.part_inst + (size_t)unit * p.T_inst * U, bytes, sm.bar_fin + which);
            }
        }
        __syncthreads();
        return *s_next;
    }
    __device__ __forceinline__ void start() { cur = grab_and_copy(0); }
    __device__ __forceinline__ void run() {
        const int N = p.N, D = p.D;
        const int th = (p.T_inst + FIN_PARTS - 1) / FIN_PARTS;
        while (cur < n_units) {
            nxt = grab_and_copy(k ^ 1);
            mbar_wait_sleepy(sm.bar_fin + k, k ? ph1 : ph0, 32);
            if (k) ph1 ^= 1; else ph0 ^= 1;
            const uint4* data = reinterpret_cast<const uint4*>(buffer(k));
            float4* comb = reinterpret_cast<float4*>(buffer(k) + (size_t)p.T_inst * U * 16);     // [FIN_PARTS][2][U] x 16 bytes
            for (int u = tid; u < FIN_PARTS * U; u += F_THREADS) {
                const int i = u % U, part = u / U;
                const int tlo = part * th, thi = min(p.T_inst, tlo + th);
                float acc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
                for (int t = tlo; t < thi; ++t) {
                    const uint4 x = data[(size_t)t * U + i];
                    const uint32_t r[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[2 * q] += __uint_as_float(r[q] << 16);
                        acc[2 * q + 1] += __uint_as_float(r[q] & 0xffff0000u);
                    }
                }
                comb[(part * 2 + 0) * U + i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                comb[(part * 2 + 1) * U + i] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
            __syncthreads();
            for (int i = tid; i < U; i += F_THREADS) {
                const int win = cur / upw;
                const int m = (cur - win * upw) * U + i, g = m >> 7, row = win * 128 + (m & 127);
                const int mt = g / (Dp / 8), c8 = g % (Dp / 8);
                if (row < N && c8 * 8 < D) {
                    float4 a = comb[i], c = comb[U + i];
#pragma unroll
                    for (int q = 1; q < FIN_PARTS; ++q) {
                        const float4 x = comb[(q * 2) * U + i], y = comb[(q * 2 + 1) * U + i];
                        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
                        c.x += y.x; c.y += y.y; c.z += y.z; c.w += y.w;
                    }
                    const float out[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
                    st_v8(p.d_inst + ((size_t)(mt * p.NS + row) * D + c8 * 8), out);
                }
            }
            cur = nxt;
            k ^= 1;
        }
    }
};

// Instance row losses (losses.py:26-39 with label smoothing) on spare CTA `hs` of `H`: the sums of z and z_y over all tiles are
// 176 KB that only the loss value needs -- on tile 0 they made that CTA the straggler of the whole grid.  One warp per row,
// lanes over the tiles (at most 160), fixed shuffle order.
__device__ __forceinline__ void spare_rows_program(const FP& p, int hs, int H, int win) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.T_inst;
    const int r0 = win * 128, Nw = min(128, p.N - r0);
    const float* ls = p.ls_inst + (size_t)win * tiles * 256;
    const float2* zz = p.zz_inst + (size_t)win * tiles * 256;
#pragma unroll 1
    for (int r = hs * 8 + w; r < 2 * Nw; r += H * 8) {
        const int mod = r / Nw, n = r - mod * Nw, srow = mod * 128 + n;
        float l[5];
        float2 z[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int t = lane + 32 * i;
            l[i] = t < tiles ? fmaxf(__ldcg(ls + (size_t)t * 256 + srow), F_NEG) : F_NEG;
            z[i] = t < tiles ? __ldcg(zz + (size_t)t * 256 + srow) : make_float2(0.f, 0.f);
        }
        float M = fmaxf(fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3])), l[4]);
        M = warp_max(M);
        float S = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            S += (lane + 32 * i < tiles) ? ex2(l[i] - M) : 0.f;
            sx += z[i].x; sy += z[i].y;
        }
        S = warp_sum(S); sx = warp_sum(sx); sy = warp_sum(sy);
        if (lane == 0) {
            const int64_t lab = p.labels[r0 + n];
            const bool bad_label = lab < 0 || lab >= (int64_t)p.C;       // the reference raises; here the row's loss becomes NaN
            const float l2 = M + log2f(S);
            p.rows_inst[mod * p.NS + r0 + n] = bad_label ? CUDART_NAN_F : l2 * F_LN2 - (1.0f - p.eps) * sy - (p.eps / (float)p.C) * sx;
        }
    }
}

// after the second grid barrier: every CTA takes an equal share of the fixed-order partial reductions
// `fi` of `G` finishing CTAs
__device__ __forceinline__ void finish_phase(const FP& p, const Smem& sm, int fi, int G) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int N = p.N, D = p.D, Dp = p.KC * 64;
    // the last CTA only forms the loss scalars; the partial reductions are shared by the others
    const int GW = p.reduce_losses && G > 1 ? G - 1 : G;
    const bool worker = fi < GW;
    PartialReducer red(p, sm);
    if (p.want_grad && p.n_inst) red.start();             // first unit in flight under the InfoNCE rows below
    F_STAMP(12);
    if (p.want_grad && p.n_nce) {      // (no block-level synchronisation below: idle CTAs / warps simply fall through)
        // d_nce[row] = normalise-backward( sum_tiles dq_part + dpos * key )   one warp per row, rows dealt round-robin to CTAs;
        // lane = 8-column chunk
        for (int row = fi + GW * (7 - w); worker && row < 2 * N; row += GW * 8) {
            const int mod = row / N, nn = row % N;
            const int grow = mod * p.NS + nn;                     // position in the full-batch arrays
            const float* key = p.key_n[mod] + (int64_t)nn * D;
            const float* qr = p.qn + (int64_t)grow * D;
            const bool on = lane * 8 < D;
            // every load of the row is issued before the first use: one L2 round trip, not four dependent ones
            const float dp = __ldcg(p.dpos + grow);
            const float inv = __ldcg(p.inv_q + grow);
            float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, q0 = k0, q1 = k0;
            if (on) {
                k0 = __ldcg(reinterpret_cast<const float4*>(key + lane * 8)); k1 = __ldcg(reinterpret_cast<const float4*>(key + lane * 8 + 4));
                q0 = __ldcg(reinterpret_cast<const float4*>(qr + lane * 8)); q1 = __ldcg(reinterpret_cast<const float4*>(qr + lane * 8 + 4));
            }
            float g[8], qv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = qv[e] = 0.f;
            float dot = 0.f;
            if (on) {
                const uint4* src = p.part_nce + ((size_t)(nn >> 7) * 2 * p.T_k + (size_t)(mod * p.T_k)) * (Dp / 8) * 128 +
                                   (size_t)(nn & 127) * (Dp / 8) + lane;                // [window][modality][tile][row][chunk]
                for (int t0 = 0; t0 < p.T_k; t0 += 16) {
                    uint4 x[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] = __ldcg(src + (size_t)min(t0 + i, p.T_k - 1) * (Dp / 8) * 128);
                    asm volatile("" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (t0 + i < p.T_k) {
                            const uint32_t r[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                g[2 * k] += __uint_as_float(r[k] << 16);
                                g[2 * k + 1] += __uint_as_float(r[k] & 0xffff0000u);
                            }
                        }
                    }
                }
                const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
                const float qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    qv[e] = qq[e];
                    g[e] = fmaf(dp, kk[e], g[e]);
                    dot = fmaf(g[e], qq[e], dot);
                }
            }
            dot = warp_sum(dot);
            if (on) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (g[e] - dot * qv[e]) * inv;
                st_v8(p.d_nce + (int64_t)grow * D + lane * 8, o);
            }
        }
    }
    if (p.want_grad && p.n_ga > 1) {
        // global-align gradient of the TEXT rows for batches above 128 rows: dq_t[row] = sum over the image windows of the block
        // contributions the align CTAs left (fixed order), then normalise backward against the bf16-rounded normalised row (the
        // value the MMAs saw).  One warp per row, taken from the low warp numbers (the InfoNCE rows above start at warp 7).
        const int NW = p.n_ga;
        for (int row = fi + GW * w; worker && row < N; row += GW * 8) {
            const int jw = row >> 7, nl = row & 127;
            const bool on = lane * 8 < D;
            float g[8], en[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = en[e] = 0.f;
            const float inv = __ldcg(p.inv_e + N + row);
            if (on) {
                const float4* er = reinterpret_cast<const float4*>(p.en + (size_t)(N + row) * D + lane * 8);
                const float4 e0 = __ldcg(er), e1 = __ldcg(er + 1);
                float4 x[8][2];
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2) {
                    const float4* src = reinterpret_cast<const float4*>(p.ga_part + ((size_t)(min(i2, NW - 1) * NW + jw) * 128 + nl) * Dp + lane * 8);
                    x[i2][0] = __ldcg(src); x[i2][1] = __ldcg(src + 1);
                }
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2) {
                    if (i2 < NW) {
                        g[0] += x[i2][0].x; g[1] += x[i2][0].y; g[2] += x[i2][0].z; g[3] += x[i2][0].w;
                        g[4] += x[i2][1].x; g[5] += x[i2][1].y; g[6] += x[i2][1].z; g[7] += x[i2][1].w;
                    }
                }
                const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) en[e] = __bfloat162float(__float2bfloat16_rn(ee[e]));
            }
            float dot = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) dot = fmaf(g[e], en[e], dot);
            dot = warp_sum(dot);
            if (on) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (g[e] - dot * en[e]) * inv;
                st_v8(p.d_ga + (size_t)(N + row) * D + lane * 8, o);
            }
        }
    }
    F_STAMP(13);
    if (p.want_grad && p.n_inst) red.run();
    if (p.reduce_losses && fi == G - 1) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int i = tid; i < 2 * N; i += F_THREADS) { a += __ldcg(p.rows_inst + i); b += __ldcg(p.rows_nce + i); }
        for (int i = tid; i < N; i += F_THREADS) c += __ldcg(p.rows_ga + i);
        a = block_sum(a, sm.red32); b = block_sum(b, sm.red32); c = block_sum(c, sm.red32);
        if (tid == 0) {
            p.losses[0] = a / (float)N;
            p.losses[1] = b / (float)N;
            p.losses[2] = c * 2.0f / (float)N;
        }
    }
}

// head.py:104-107 for CTA `part` of `parts`: queue[:, ptr:ptr+N] = keys^T (both modalities) and id_queue[ptr:ptr+N] = ids.
// Element order [modality][d][n] with n fastest: the queue writes of a warp are contiguous.  The pointer is read, never written,
// here (it moves after the grid barrier); a pointer outside [0, K-N] (a checkpoint taken with another batch size) wraps instead
// of writing past the row / the allocation.
__device__ __forceinline__ void enqueue_slice(const FP& p, int part, int parts) {
    const int N = p.N, D = p.D, K = p.K;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = (int)(((*p.enq_ptr % K) + K) % K);        // one 64-bit division per thread; 32-bit arithmetic below
    // one warp per queue row (modality, d): lanes sweep the N batch columns, four loads in flight per lane
#pragma unroll 1
    for (int r = part * 8 + w; r < 2 * D; r += parts * 8) {
        const int mod = r >= D, d = r - mod * D;
        const float* src = (mod ? p.key_n[0] : p.key_n[1]) + d;                     // key_n[0] = t_key_n, key_n[1] = v_key_n
        float* dst = p.enq_queue[mod] + (int64_t)d * K;
#pragma unroll 1
        for (int n0 = lane; n0 < N; n0 += 128) {
            float x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = __ldcg(src + (int64_t)min(n0 + 32 * i, N - 1) * D);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = n0 + 32 * i;
                int col = p0 + n;
                col -= col >= K ? K : 0;
                if (n < N) dst[col] = x[i];
            }
        }
    }
    if (part == parts - 1)
        for (int n = threadIdx.x; n < N; n += F_THREADS) {
            int col = p0 + n;
            col -= col >= K ? K : 0;
            p.enq_ids[col] = p.labels[n];
        }
}

__global__ void __launch_bounds__(F_THREADS, 1) fused_loss_kernel(const FP p) {
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic (not through an integer) so that the compiler keeps the shared address space: LDS/STS, not LD/ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    Smem sm;
    sm.E = smem + F_OFF_E; sm.DZ = smem + F_OFF_DZ; sm.WB = smem + F_OFF_WB;
    uint8_t* misc = smem + F_OFF_MISC;
    sm.col = reinterpret_cast<float2*>(misc);
    sm.bar_load = reinterpret_cast<uint64_t*>(misc + 1024);
    sm.bar_mma = reinterpret_cast<uint64_t*>(misc + 1032);
    sm.bar_dw = reinterpret_cast<uint64_t*>(misc + 1040);
    sm.tmem_slot = reinterpret_cast<uint32_t*>(misc + 1048);
    sm.bar_fin = reinterpret_cast<uint64_t*>(misc + 1056);
    sm.red32 = reinterpret_cast<float*>(misc + 1088);
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        mbar_init(sm.bar_load, 1);
        mbar_init(sm.bar_mma, 1);
        mbar_init(sm.bar_dw, 1);
        mbar_init(sm.bar_fin, 1);
        mbar_init(sm.bar_fin + 1, 1);
        mbar_fence_init();
    }
    // CTA roles by block index: instance tiles, InfoNCE tiles, the align CTA, then SPARE CTAs (the rest of the SMs) that own no
    // tile and only take part in the reductions at the end
    const int b = blockIdx.x, G = (int)gridDim.x;
    const int n_tiles = p.n_inst + p.n_nce + p.n_ga;
    const bool has_tile = b < n_tiles;
    if (warp == 0 && has_tile) tmem_alloc(sm.tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = has_tile ? *sm.tmem_slot : 0u;
    F_STAMP(0);

    uint32_t mma_phase = 0, load_phase = 0;
    if (b < p.n_inst) tile_program<true>(p, sm, tmem, 0, b, mma_phase, load_phase);
    else if (b < p.n_inst + p.n_nce) {
        if (p.nce_dual) {                // one CTA per queue tile index: the image queries' tile, then the text queries' tile
#pragma unroll 1
            for (int mod = 0; mod < 2; ++mod) {
                tile_program<false>(p, sm, tmem, mod, b - p.n_inst, mma_phase, load_phase);
                __syncthreads();         // the shared scratch of the first pass is re-used by the second
            }
        } else {
            tile_program<false>(p, sm, tmem, (b - p.n_inst) / p.T_k, (b - p.n_inst) % p.T_k, mma_phase, load_phase);
        }
    } else if (has_tile) align_program(p, sm, tmem, b - p.n_inst - p.n_nce);
    else griddep_wait();                 // spare CTAs: the counters they poll are cleared by the prologue launch

    // Second counter: every tile CTA arrives once its partial dE tiles (and enqueue slice) are written.  With `fin_early` the
    // instance tiles -- the longest chain of the kernel -- arrived before their dW epilogue (tile_program), the align CTA after
    // its row losses (align_program), and both are done here: the fixed-order reductions of the partial tiles belong to the
    // InfoNCE CTAs (which finish early) and the spare CTAs, so they run UNDER the dW writes instead of after them.
    // Otherwise every CTA reduces a share.
    const bool is_nce = b >= p.n_inst && b < p.n_inst + p.n_nce;
    // _dequeue_and_enqueue: the queues were last read by the prologue's re-pack, so the CTAs with time to spare write the key
    // columns -- the spare CTAs (idle until the reductions) when there are any, else the InfoNCE / align CTAs after their tiles
    const int n_spare = G - n_tiles;
    if (p.enq_ptr != nullptr) {
        if (n_spare > 0) { if (!has_tile) enqueue_slice(p, b - n_tiles, n_spare); }
        else if (is_nce || (has_tile && b >= p.n_inst && !p.fin_early)) enqueue_slice(p, b - p.n_inst, p.n_nce + (p.fin_early ? 0 : p.n_ga));
    }
    const bool finisher = !p.fin_early || is_nce || !has_tile;
    if (!has_tile) {                                         // spare CTAs: instance row losses once every tile's statistics are out
        if (b - n_tiles < p.row_helpers) {
            for (int win = 0; win < (p.N + 127) / 128; ++win) {
                grid_wait(p.bar, (unsigned)(p.n_inst * (win + 1)));
                spare_rows_program(p, b - n_tiles, p.row_helpers, win);
            }
        }
        grid_arrive(p.bar + 1);                              // (their enqueue slices and row losses are out)
    }
    if (has_tile && (!p.fin_early || is_nce)) grid_arrive(p.bar + 1);
    if (finisher) {
        const int fi = !p.fin_early ? b : (is_nce ? b - p.n_inst : b - p.n_inst - p.n_ga);
        const int nf = !p.fin_early ? G : G - p.n_inst - p.n_ga;
        grid_wait(p.bar + 1, (unsigned)G);                   // every tile CTA and every spare CTA arrives once
        F_STAMP(8);
        // every enqueue slice has read the old pointer before it arrived: head.py:108-109
        if (p.enq_ptr != nullptr && fi == nf - 1 && threadIdx.x == 0) *p.enq_ptr = (*p.enq_ptr + p.N) % p.K;
        finish_phase(p, sm, fi, nf);
        F_STAMP(9);
    } else if (p.want_grad) {
        // an instance tile that is done with its dW epilogue: take whatever units of the partial-tile reduction are left
        grid_wait(p.bar + 1, (unsigned)G);
        PartialReducer red(p, sm);
        red.start();
        red.run();
        F_STAMP(9);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0 && has_tile) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

struct Scratch {
    uint8_t *Ep, *ENp, *QNp, *QUp;
    float *ls_inst, *ls_nce;
    float2* zz_inst;
    uint4 *part_inst, *part_nce;
    float* ga_part;
    unsigned* bar;
    unsigned long long* dbg;       // [160][16] phase timestamps (TRB_FUSED_DEBUG)
    float* dbg_logits;             // [256][128] logits of one instance tile (TRB_FUSED_DEBUG_LOGITS)
    int64_t bytes;
};

Scratch carve_scratch(uint8_t* base, int N, int D, int K, int C) {
    const int64_t NW = (N + 127) / 128;               // per-window buffers: operand images, statistics, partial tiles
    const int KC = (D + 127) / 128 * 2, Dp = KC * 64;
    const int T_inst = (C + F_TILE - 1) / F_TILE, T_k = (K + F_TILE - 1) / F_TILE;
    Scratch s;
    uint8_t* p = base;
    auto take = [&](int64_t bytes) { uint8_t* r = p; p += (bytes + 1023) / 1024 * 1024; return r; };
    const int64_t img = NW * 2 * KC * BLOCK_BYTES;
    s.Ep = take(img); s.ENp = take(img); s.QNp = take(img);
    s.QUp = take((int64_t)2 * T_k * F_WB_BYTES);
    s.ls_inst = reinterpret_cast<float*>(take(NW * 256 * T_inst * 4));
    s.zz_inst = reinterpret_cast<float2*>(take(NW * 256 * T_inst * 8));
    s.ls_nce = reinterpret_cast<float*>(take(NW * 256 * T_k * 4));
    s.part_inst = reinterpret_cast<uint4*>(take(NW * T_inst * 256 * Dp * 2));
    s.part_nce = reinterpret_cast<uint4*>(take(NW * 2 * T_k * 128 * Dp * 2));
    s.ga_part = reinterpret_cast<float*>(take(NW > 1 ? NW * NW * 128 * Dp * 4 : 0));
    s.bar = reinterpret_cast<unsigned*>(take(256));
    s.dbg = reinterpret_cast<unsigned long long*>(take(160 * 16 * 8));
    s.dbg_logits = reinterpret_cast<float*>(take(256 * 128 * 4));
    s.bytes = p - base;
    return s;
}

}  // namespace

// debug read-backs (declared in include/textreid_b200.h): the buffers live in the CALLER's workspace, the library owns nothing
int fused_loss_debug_copy(const uint8_t* scratch, int N, int D, int K, int C, int what, void* host_out) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~uintptr_t(1023));
    const Scratch s = carve_scratch(base, N, D, K, C);
    if (what == 0) return (int)cudaMemcpy(host_out, s.dbg, 160 * 16 * 8, cudaMemcpyDeviceToHost);
    return (int)cudaMemcpy(host_out, s.dbg_logits, 256 * 128 * 4, cudaMemcpyDeviceToHost);
}

// everything in one cooperative launch (2 launches per step with the prologue)
bool fused_loss_supported(int N, int D, int K, int C, int sm_count) {
    if (N < 1 || N > 256 || D < 64 || D > 256 || (D % 64) != 0) return false;
    const int T_inst = (C + F_TILE - 1) / F_TILE, T_k = (K + F_TILE - 1) / F_TILE;
    // up to 128 rows: one CTA per InfoNCE tile and modality, one align CTA; up to 256 rows: an InfoNCE CTA takes both modalities
    // of its tile in turn, one align CTA per 128-row window
    const int ctas = N <= 128 ? T_inst + 2 * T_k + 1 : T_inst + T_k + 2;
    return ctas <= sm_count;
}

// instance + InfoNCE branches in one cooperative launch, global-align on the unfused sequence: batches of up to 1024 rows
bool fused_windows_supported(int N, int D, int K, int C, int sm_count) {
    if (N <= 128 || N > 1024 || D < 64 || D > 256 || (D % 64) != 0 || (D % 8) != 0) return false;
    return (C + F_TILE - 1) / F_TILE + (K + F_TILE - 1) / F_TILE <= sm_count;
}

int64_t fused_loss_scratch_bytes(int N, int D, int K, int C) { return carve_scratch(nullptr, N, D, K, C).bytes + 1024; }

static ProArgs make_pro_args(const FusedLossArgs& a, const Scratch& s) {
    ProArgs q;
    q.v_embed = a.v_embed; q.t_embed = a.t_embed; q.v_qraw = a.v_qraw; q.t_qraw = a.t_qraw; q.v_key = a.v_key; q.t_key = a.t_key;
    q.v_queue = a.v_queue; q.t_queue = a.t_queue;
    q.v_key_n = a.v_key_n; q.t_key_n = a.t_key_n; q.E2 = a.E2; q.en = a.en; q.inv_e = a.inv_e; q.qn = a.qn; q.inv_q = a.inv_q; q.pos = a.pos;
    q.Ep = s.Ep; q.ENp = s.ENp; q.QNp = s.QNp; q.QUp = s.QUp;
    q.normalize_keys = a.normalize_keys; q.N = a.N; q.D = a.D; q.KC = (a.D + 127) / 128 * 2; q.K = a.K;
    q.T_k = (a.K + F_TILE - 1) / F_TILE;
    q.W = ((a.roles & 1) && !getenv("TRB_FUSED_NO_PREFETCH")) ? a.projection : nullptr;
    q.W_bytes = (int64_t)a.D * a.C * 4;
    return q;
}

int fused_loss_prologue(const FusedLossArgs& a, cudaStream_t st) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a.scratch) + 1023) & ~uintptr_t(1023));
    const Scratch s = carve_scratch(base, a.N, a.D, a.K, a.C);
    static TrbDeviceOnce attr;
    if (trb_first_on_device(attr))   // same shared-memory carve-out as the cooperative kernel that follows: no SM reconfiguration between the two
        TRB_CUDA_OK(cudaFuncSetAttribute(fused_prologue_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const ProArgs q = make_pro_args(a, s);
    const int ntasks = 256 * ((a.N + 127) / 128) + ((a.roles & 2) ? 2 * q.KC * 64 : 0);   // queue re-pack only for fused InfoNCE tiles
    fused_prologue_kernel<<<(ntasks + 7) / 8, 256, 0, st>>>(q, s.bar, ntasks);
    TRB_LAUNCH_OK();
    return 0;
}

int fused_loss_launch(const FusedLossArgs& a, cudaStream_t st) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a.scratch) + 1023) & ~uintptr_t(1023));
    const Scratch s = carve_scratch(base, a.N, a.D, a.K, a.C);
    FP p;
    memset(&p, 0, sizeof(p));
    p.N = a.N; p.D = a.D; p.K = a.K; p.C = a.C;
    p.KC = (a.D + 127) / 128 * 2;
    p.NS = a.N;
    p.Nn = a.N;
    p.T_inst = (a.C + F_TILE - 1) / F_TILE; p.T_k = (a.K + F_TILE - 1) / F_TILE;
    p.roles = a.roles;
    p.n_inst = (a.roles & 1) ? p.T_inst : 0;
    p.nce_dual = (a.roles & 2) && a.N > 128 ? 1 : 0;
    p.n_nce = (a.roles & 2) ? (p.nce_dual ? p.T_k : 2 * p.T_k) : 0;
    p.n_ga = (a.roles & 4) ? (a.N + 127) / 128 : 0;      // one align CTA per 128-row window of the image rows
    p.want_grad = a.d_inst != nullptr;
    p.reduce_losses = a.reduce_losses;
    p.T = a.T; p.eps = a.eps; p.alpha = a.alpha; p.beta = a.beta; p.sp = a.sp; p.sn = a.sn;
    p.W = a.projection;
    p.queue[0] = a.t_queue; p.queue[1] = a.v_queue;
    p.key_n[0] = a.t_key_n; p.key_n[1] = a.v_key_n;
    p.labels = a.labels; p.id_queue = a.id_queue;
    p.Ep = s.Ep; p.ENp = s.ENp; p.QNp = s.QNp; p.QUp = s.QUp;
    p.en = a.en; p.qn = a.qn; p.inv_e = a.inv_e; p.inv_q = a.inv_q; p.pos = a.pos;
    p.ls_inst = s.ls_inst; p.zz_inst = s.zz_inst; p.ls_nce = s.ls_nce;
    // debug only (both live in the caller's workspace): phase timestamps of every CTA / the logits of one instance tile
    p.dbg = getenv("TRB_FUSED_DEBUG") ? s.dbg : nullptr;
    p.dbg_logits = nullptr;
    if (const char* e = getenv("TRB_FUSED_DEBUG_LOGITS")) { p.dbg_logits = s.dbg_logits; p.dbg_tile = atoi(e); }
    p.enq_queue[0] = a.enq_v_queue; p.enq_queue[1] = a.enq_t_queue; p.enq_ids = a.enq_ids; p.enq_ptr = a.enq_ptr;
    p.part_inst = s.part_inst; p.part_nce = s.part_nce; p.ga_part = s.ga_part;
    p.dpos = a.dpos; p.rows_inst = a.rows_inst; p.rows_nce = a.rows_nce; p.rows_ga = a.rows_ga;
    p.losses = a.losses; p.d_inst = a.d_inst; p.d_nce = a.d_nce; p.d_ga = a.d_ga; p.d_proj = a.d_proj;
    p.bar = s.bar;
    const int n_tiles = p.n_inst + p.n_nce + p.n_ga;
    if (n_tiles == 0) return 0;
    // one CTA per SM: the SMs without a tile get a spare CTA that only reduces partial tiles at the end
    int dev = 0, sms = 0;
    TRB_CUDA_OK(cudaGetDevice(&dev));
    TRB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // spare CTAs: all remaining SMs when the whole step is this one launch; a launch that shares the device with the unfused
    // branches on the helper streams (batches above 128 rows, partial roles) keeps 16+ SMs free for them -- a cooperative grid that covers
    // every SM would have to wait until those kernels have drained; without instance tiles spare CTAs have nothing to do
    int spare = sms - n_tiles;
    if (spare < 0 || p.n_inst == 0) spare = 0;
    if (a.roles != 7) spare = spare - 16 > 32 ? 32 : (spare > 16 ? spare - 16 : 0);
    const int grid = n_tiles + spare;
    p.fin_early = (p.n_inst > 0 && p.want_grad && grid - p.n_inst - p.n_ga >= 24) ? 1 : 0;
    {
        const int cap = (F_OFF_MISC / 2) / (16 * p.T_inst + 32 * FIN_PARTS);     // per buffer: data + combine scratch
        p.fin_U = cap >= 64 ? 64 : (cap >= 32 ? 32 : (cap >= 16 ? 16 : 8));
    }
    p.row_helpers = p.n_inst > 0 ? (grid - n_tiles < 16 ? grid - n_tiles : 16) : 0;

    static TrbDeviceOnce attr;
    if (trb_first_on_device(attr))
        TRB_CUDA_OK(cudaFuncSetAttribute(fused_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    // all CTAs meet at two grid barriers: the launch must be co-resident (cooperative), one CTA per SM
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(F_THREADS);
    cfg.dynamicSmemBytes = F_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    // programmatic dependent launch behind the prologue kernel (TRB_FUSED_PDL=0 turns it off)
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    const char* pdl = getenv("TRB_FUSED_PDL");
    cfg.numAttrs = (a.after_prologue && !(pdl && atoi(pdl) == 0)) ? 2 : 1;
    TRB_CUDA_OK(cudaLaunchKernelEx(&cfg, fused_loss_kernel, p));
    return 0;
}
