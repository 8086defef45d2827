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
constexpr int F_TILE = 128;                         // classes / queue slots per CTA
constexpr int F_E_BYTES = 8 * BLOCK_BYTES;          // 128 KiB  embedding rows: [2 row-blocks][<=4 k-chunks][16 KiB]
constexpr int F_DZ_BYTES = 2 * BLOCK_BYTES;         // 32 KiB   logit gradient of one 128-row block: [2 column chunks][16 KiB]
constexpr int F_WB_BYTES = 2 * 256 * 128;           // 64 KiB   W / queue tile: [2 column chunks][256 d][128 B]
constexpr int F_WB_CHUNK = 256 * 128;
constexpr int F_OFF_E = 0, F_OFF_DZ = F_E_BYTES, F_OFF_WB = F_OFF_DZ + F_DZ_BYTES, F_OFF_MISC = F_OFF_WB + F_WB_BYTES;
constexpr int F_MISC_BYTES = 1024;
constexpr int F_SMEM = F_OFF_MISC + F_MISC_BYTES + 1024;   // + alignment slack

struct FP {
    int N, D, K, C, KC, T_inst, T_k, n_inst, n_nce, n_ga, want_grad, reduce_losses, roles, variant;
    float T, eps, alpha, beta, sp, sn;
    const float* W;
    const float* queue[2];          // queue scored by modality m's queries: [0] = t_queue, [1] = v_queue  (head.py:162,168)
    const float* key_n[2];          // positive key of modality m: [0] = t_key_n, [1] = v_key_n            (head.py:160,166)
    const int64_t *labels, *id_queue;
    const uint8_t *Ep, *ENp, *QNp;
    const float *en, *qn, *inv_e, *inv_q, *pos;
    float4 *st_inst, *st_nce;
    float *part_inst, *part_nce, *dpos, *rows_inst, *rows_nce, *rows_ga, *losses, *d_inst, *d_nce, *d_ga, *d_proj;
    unsigned* bar;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MN-major SWIZZLE_128B descriptor: 64-element groups along M/N are `lbo` bytes apart, 8-row groups along K 1024 B apart.
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo, int variant) {
    uint32_t l = lbo >> 4, s = 1024 >> 4;
    if (variant & 1) { const uint32_t t = l; l = s; s = t; }      // debug: swapped interpretation
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(l & 0x3FFF) << 16;
    d |= (uint64_t)(s & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return umma_idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned total) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        const long long t0 = clock64();
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= total) break;
            __nanosleep(40);
            if (clock64() - t0 > 4000000000LL) {
                printf("trb: fused loss grid barrier timed out (block %d, %u of %u)\n", blockIdx.x, v, total);
                __trap();
            }
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

// fp32 tile src[d, c0 + c] (d < nrows, c < 128) -> bf16 image in shared memory; per-column sums of squares -> red[8][128]
__device__ __forceinline__ void load_tile_bf16(const float* __restrict__ src, int64_t ld, int nrows, int rows_pad, int c0,
                                               int ncols, uint8_t* wb, float* red) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
    bool cv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cv[j] = c0 + lane + 32 * j < ncols;
    for (int i0 = 0; i0 < rows_pad / 8; i0 += 4) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = w + 8 * (i0 + u);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[u][j] = (d < nrows && cv[j]) ? __ldg(src + (int64_t)d * ld + c0 + lane + 32 * j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = w + 8 * (i0 + u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ss[j] = fmaf(v[u][j], v[u][j], ss[j]);
                *reinterpret_cast<__nv_bfloat16*>(wb + wb_offset(d, lane + 32 * j)) = __float2bfloat16_rn(v[u][j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[w * 128 + lane + 32 * j] = ss[j];
}

struct Smem {
    uint8_t *E, *DZ, *WB;
    float* inv;            // [128] column scale: 1/||w_c|| (instance) or 1/T (InfoNCE); 0 for excluded columns
    uint32_t* valid;       // [4]   bit c: column takes part in the softmax
    uint64_t *bar_load, *bar_mma;
    uint32_t* tmem_slot;
    float* red32;          // [32]  block_sum scratch
};

// combine the per-tile softmax statistics of one row: returns lse; sz / zy = sums of the 3rd / 4th statistic
__device__ __forceinline__ float combine_stats(const float4* __restrict__ st, int tiles, float m0, float s0, float& sz, float& zy) {
    float M = m0, S = s0;
    sz = 0.f; zy = 0.f;
    for (int t = 0; t < tiles; ++t) {
        const float4 a = __ldcg(st + t);
        sz += a.z; zy += a.w;
        if (a.y > 0.f) {
            if (a.x > M) { S = S * expf(M - a.x) + a.y; M = a.x; }
            else S += a.y * expf(a.x - M);
        }
    }
    return M + logf(S);
}

// ------------------------------------------------------------------------------------------------------------------------
// instance (INST) / InfoNCE tile
// ------------------------------------------------------------------------------------------------------------------------
template <bool INST>
__device__ void tile_program(const FP& p, const Smem& sm, uint32_t tmem, int mod, int tile) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, q = w & 3, h = w >> 2;
    const int n = q * 32 + lane;                       // row inside a 128-row block = TMEM lane
    const int N = p.N, Dp = p.KC * 64;
    const int c0 = tile * F_TILE;
    constexpr int MT = INST ? 2 : 1;
    const int ncols = INST ? p.C : p.K;
    const int tiles = INST ? p.T_inst : p.T_k;
    const uint32_t lanes = (uint32_t)(q * 32) << 16;
    uint32_t mma_phase = 0;
    float* red = reinterpret_cast<float*>(sm.DZ);      // [8][128] + [128]: free whenever no logit gradient is staged

    // ---- operand rows: packed bf16 image written by the prologue (instance: raw embeds, both modalities; InfoNCE: q rows)
    if (tid == 0) {
        const uint8_t* src = INST ? p.Ep : p.QNp + (size_t)mod * p.KC * BLOCK_BYTES;
        const int blocks = MT * p.KC;
        mbar_expect_tx(sm.bar_load, (uint32_t)blocks * BLOCK_BYTES);
        for (int b = 0; b < blocks; ++b) bulk_g2s(sm.E + (size_t)b * BLOCK_BYTES, src + (size_t)b * BLOCK_BYTES, BLOCK_BYTES, sm.bar_load);
    }

    // ---- W / queue tile: HBM -> bf16 shared image, column statistics
    load_tile_bf16(INST ? p.W : p.queue[mod], ncols, p.D, Dp, c0, ncols, sm.WB, red);
    __syncthreads();
    if (tid < 128) {
        float tot = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) tot += red[ww * 128 + tid];
        bool ok = c0 + tid < ncols;
        float inv = 0.f;
        if (INST) {
            if (ok) inv = __fdiv_rn(1.0f, fmaxf(sqrtf(tot), 1e-12f));           // losses.py:51
        } else {
            if (ok) {                                                           // head.py:148-157: drop slots holding a batch id
                const int64_t id = p.id_queue[c0 + tid];
                bool hit = false;
                for (int i = 0; i < N; ++i) hit |= (p.labels[i] == id);
                ok = !hit;
            }
            if (ok) inv = __fdiv_rn(1.0f, p.T);
        }
        sm.inv[tid] = inv;
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) sm.valid[w] = bal;
    }
    fence_async_smem();
    __syncthreads();

    // ---- forward logits: z[mt] = E[mt] . Wb   (M = 128 rows, N = 128 columns, K = Dp)
    if (tid == 0) {
        mbar_wait(sm.bar_load, 0);
        tc_fence_after();
        const uint32_t id_f = idesc(128, 128, 0, 1);
        const uint32_t e0 = smem_u32(sm.E), wb0 = smem_u32(sm.WB);
        for (int mt = 0; mt < MT; ++mt)
            for (int ks = 0; ks < Dp / 16; ++ks)
                umma_bf16(tmem + mt * 128, umma_desc_sw128(e0 + (mt * p.KC + (ks >> 2)) * BLOCK_BYTES + (ks & 3) * 32),
                          desc_mn(wb0 + ks * 2048, F_WB_CHUNK, p.variant), id_f, (uint32_t)(ks > 0));
        umma_commit(sm.bar_mma);
    }
    mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
    mma_phase ^= 1;
    tc_fence_after();

    // ---- per-row partial softmax statistics of this tile (thread = row n of block h)
    if (h < MT) {
        const int y = INST ? (int)p.labels[n < N ? n : 0] - c0 : -1;
        float m = -CUDART_INF_F, s = 0.f, sz = 0.f, zy = 0.f;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float v[32];
            tmem_ld32(tmem + lanes + (uint32_t)(h * 128 + j * 32), v);
            const uint32_t vm = sm.valid[j];
            float cm = -CUDART_INF_F;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float z = v[i] * sm.inv[j * 32 + i];
                const bool ok = (vm >> i) & 1u;
                sz += ok ? z : 0.f;
                if (INST && j * 32 + i == y) zy = z;
                v[i] = ok ? z : -CUDART_INF_F;
                cm = fmaxf(cm, v[i]);
            }
            if (cm > -CUDART_INF_F) {
                const float nm = fmaxf(m, cm);
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += expf(v[i] - nm);
                s = s * expf(m - nm) + acc;
                m = nm;
            }
        }
        if (n < N) {
            float4* st = (INST ? p.st_inst : p.st_nce) + (size_t)((INST ? h : mod) * 128 + n) * tiles + tile;
            *st = make_float4(m, s, sz, zy);
        }
    }

    grid_barrier(p.bar, gridDim.x);

    // ---- row log-sum-exp over all tiles; the first tile of a row set also writes the row losses
    float lse[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        lse[mt] = 0.f;
        if (n < N) {
            const int rb = INST ? mt : mod;
            const float4* st = (INST ? p.st_inst : p.st_nce) + (size_t)(rb * 128 + n) * tiles;
            float sz, zy;
            if (INST) {
                lse[mt] = combine_stats(st, tiles, -CUDART_INF_F, 0.f, sz, zy);
                if (tile == 0 && h == mt)                                       // losses.py:26-39 with label smoothing
                    p.rows_inst[mt * N + n] = lse[mt] - (1.0f - p.eps) * zy - (p.eps / (float)p.C) * sz;
            } else {
                const float z0 = __fdiv_rn(p.pos[mod * N + n], p.T);            // column 0 of the reference's logits
                lse[mt] = combine_stats(st, tiles, z0, 1.0f, sz, zy);
                if (tile == 0 && h == 0) {                                      // losses.py:206-217, target 0
                    p.rows_nce[mod * N + n] = lse[mt] - z0;
                    p.dpos[mod * N + n] = (expf(z0 - lse[mt]) - 1.0f) / ((float)N * p.T);
                }
            }
        }
    }

    if (p.want_grad) {
        const float invN = 1.0f / (float)N;
        const float onehot = INST ? 1.0f - p.eps : 0.f, uni = INST ? p.eps / (float)p.C : 0.f;
        const bool do_dw = INST && p.d_proj != nullptr;
        const int NH = Dp / 128;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            // -- logit gradient of row block mt -> bf16 shared image (thread = row n, 64 columns of chunk h)
            const int y = INST ? (int)p.labels[n < N ? n : 0] - c0 : -1;
            const float lse_row = (mt == 0) ? lse[0] : lse[MT - 1];
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int j = h * 2 + jj;
                float v[32];
                tmem_ld32(tmem + lanes + (uint32_t)(mt * 128 + j * 32), v);
                const uint32_t vm = sm.valid[j];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = j * 32 + i;
                    const float sc = sm.inv[c];
                    const float z = v[i] * sc;
                    const float g = expf(z - lse_row) - ((c == y) ? onehot : 0.f) - uni;
                    const bool ok = ((vm >> i) & 1u) && (n < N);
                    v[i] = ok ? g * invN * sc : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint4 o;
                    o.x = pack2(v[8 * k + 0], v[8 * k + 1]); o.y = pack2(v[8 * k + 2], v[8 * k + 3]);
                    o.z = pack2(v[8 * k + 4], v[8 * k + 5]); o.w = pack2(v[8 * k + 6], v[8 * k + 7]);
                    *reinterpret_cast<uint4*>(sm.DZ + dz_offset(h, n, jj * 4 + k)) = o;
                }
            }
            tc_fence_before();
            fence_async_smem();
            __syncthreads();

            const int rounds = INST ? NH : 1;          // instance: dE in 128-column rounds that reuse z[mt]'s TMEM columns
            for (int r = 0; r < rounds; ++r) {
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t e0 = smem_u32(sm.E), wb0 = smem_u32(sm.WB), dz0 = smem_u32(sm.DZ);
                    if (do_dw && r == 0) {
                        // dWs[d, c] += sum_rows E[row, d] * dz'[row, c]   (M = 128 d per half, N = 128 columns, K = 128 rows)
                        const uint32_t id_w = idesc(128, 128, 1, 1);
                        for (int hh = 0; hh < NH; ++hh)
                            for (int ks = 0; ks < 8; ++ks)
                                umma_bf16(tmem + 256 + hh * 128,
                                          desc_mn(e0 + (mt * p.KC + 2 * hh) * BLOCK_BYTES + ks * 2048, BLOCK_BYTES, p.variant),
                                          desc_mn(dz0 + ks * 2048, BLOCK_BYTES, p.variant), id_w, (uint32_t)((mt > 0) | (ks > 0)));
                    }
                    // dE[row, d] = sum_c dz'[row, c] * Wb[d, c]           (M = 128 rows, N = d, K = 128 columns)
                    const int nd = INST ? 128 : Dp;
                    const uint32_t id_e = idesc(128, nd, 0, 0);
                    const uint32_t dcol = INST ? (uint32_t)(mt * 128) : 256u;
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem + dcol, umma_desc_sw128(dz0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32),
                                  umma_desc_sw128(wb0 + (ks >> 2) * F_WB_CHUNK + r * (128 * 128) + (ks & 3) * 32), id_e,
                                  (uint32_t)(ks > 0));
                    umma_commit(sm.bar_mma);
                }
                mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
                mma_phase ^= 1;
                tc_fence_after();
                // -- drain the partial dE of this tile to global (reduced over tiles after the second grid barrier)
                if (INST) {
                    float* dst = p.part_inst + ((size_t)tile * 256 + mt * 128 + n) * Dp + r * 128 + h * 64;
#pragma unroll 1
                    for (int jj = 0; jj < 2; ++jj) {
                        float v[32];
                        tmem_ld32(tmem + lanes + (uint32_t)(mt * 128 + h * 64 + jj * 32), v);
                        if (n < N) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                __stcg(reinterpret_cast<float4*>(dst + jj * 32) + k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                        }
                    }
                } else {
                    const int half = Dp / 2;
                    float* dst = p.part_nce + (((size_t)mod * p.T_k + tile) * 128 + n) * Dp + h * half;
#pragma unroll 1
                    for (int jj = 0; jj < half / 32; ++jj) {
                        float v[32];
                        tmem_ld32(tmem + lanes + (uint32_t)(256 + h * half + jj * 32), v);
                        if (n < N) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                __stcg(reinterpret_cast<float4*>(dst + jj * 32) + k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                        }
                    }
                }
                tc_fence_before();
                __syncthreads();       // TMEM columns and (after the last round) the dz' image may be overwritten
            }
        }

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
            __syncthreads();
            bool cv[4];
            float sc[4], acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { cv[j] = c0 + lane + 32 * j < p.C; sc[j] = sm.inv[lane + 32 * j]; acc[j] = 0.f; }
            for (int d = w; d < p.D; d += 8) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = lane + 32 * j;
                    if (cv[j]) {
                        const float wv = __ldg(p.W + (int64_t)d * p.C + c0 + c) * sc[j];
                        acc[j] = fmaf(tile1[d * 128 + ((((c >> 2) ^ (d & 7))) << 2) + (c & 3)], wv, acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) red[w * 128 + lane + 32 * j] = acc[j];
            __syncthreads();
            if (tid < 128) {
                float t = 0.f;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) t += red[ww * 128 + tid];
                red[1024 + tid] = t;
            }
            __syncthreads();
            float dot[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) dot[j] = red[1024 + lane + 32 * j];
            for (int d = w; d < p.D; d += 8) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = lane + 32 * j;
                    if (cv[j]) {
                        const float wv = __ldg(p.W + (int64_t)d * p.C + c0 + c) * sc[j];
                        p.d_proj[(int64_t)d * p.C + c0 + c] = tile1[d * 128 + ((((c >> 2) ^ (d & 7))) << 2) + (c & 3)] - wv * dot[j];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// global align (losses.py:102-128): S = en_v en_t^T, pair losses, dS, dq_v = dS en_t, dq_t = dS^T en_v, normalise backward
// ------------------------------------------------------------------------------------------------------------------------
__device__ void align_program(const FP& p, const Smem& sm, uint32_t tmem) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, q = w & 3, h = w >> 2;
    const int n = q * 32 + lane;
    const int N = p.N, D = p.D, Dp = p.KC * 64;
    const uint32_t lanes = (uint32_t)(q * 32) << 16;
    uint32_t mma_phase = 0;
    int64_t* s_lab = reinterpret_cast<int64_t*>(sm.WB);                 // [128]
    float* s_part = reinterpret_cast<float*>(sm.WB + 2048);             // [2][128]

    if (tid == 0) {
        const int blocks = 2 * p.KC;
        mbar_expect_tx(sm.bar_load, (uint32_t)blocks * BLOCK_BYTES);
        for (int b = 0; b < blocks; ++b) bulk_g2s(sm.E + (size_t)b * BLOCK_BYTES, p.ENp + (size_t)b * BLOCK_BYTES, BLOCK_BYTES, sm.bar_load);
    }
    // nothing here depends on the tiles' statistics: arrive at the first grid barrier right away, never wait on it
    if (tid == 0) atomicAdd(p.bar, 1u);
    if (tid < 128) s_lab[tid] = tid < N ? p.labels[tid] : (int64_t)-1;
    __syncthreads();
    const uint32_t e0 = smem_u32(sm.E), dz0 = smem_u32(sm.DZ);
    const uint32_t et0 = e0 + p.KC * BLOCK_BYTES;                       // text rows
    if (tid == 0) {
        mbar_wait(sm.bar_load, 0);
        tc_fence_after();
        const uint32_t id_s = idesc(128, 128, 0, 0);
        for (int ks = 0; ks < Dp / 16; ++ks)
            umma_bf16(tmem, umma_desc_sw128(e0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32),
                      umma_desc_sw128(et0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32), id_s, (uint32_t)(ks > 0));
        umma_commit(sm.bar_mma);
    }
    mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
    mma_phase ^= 1;
    tc_fence_after();

    {
        const int64_t yi = s_lab[n];
        const float two_over_n = 2.0f / (float)N;
        float acc = 0.f;
#pragma unroll 1
        for (int jj = 0; jj < 2; ++jj) {
            const int j = h * 2 + jj;
            float v[32];
            tmem_ld32(tmem + lanes + (uint32_t)(j * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int col = j * 32 + i;
                const bool ok = n < N && col < N;
                const bool same = s_lab[col] == yi;
                const float x = same ? -p.sp * (v[i] - p.alpha) : p.sn * (v[i] - p.beta);
                const float e = expf(x);
                acc += ok ? logf(1.0f + e) : 0.f;                       // the reference's literal log(1+exp(x)), losses.py:123-124
                v[i] = ok ? (same ? -p.sp : p.sn) * (e / (1.0f + e)) * two_over_n : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint4 o;
                o.x = pack2(v[8 * k + 0], v[8 * k + 1]); o.y = pack2(v[8 * k + 2], v[8 * k + 3]);
                o.z = pack2(v[8 * k + 4], v[8 * k + 5]); o.w = pack2(v[8 * k + 6], v[8 * k + 7]);
                *reinterpret_cast<uint4*>(sm.DZ + dz_offset(h, n, jj * 4 + k)) = o;
            }
        }
        s_part[h * 128 + n] = acc;
    }
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    if (h == 0 && n < N) p.rows_ga[n] = s_part[n] + s_part[128 + n];

    if (p.want_grad) {
        if (tid == 0) {
            tc_fence_after();
            // dq_v[i, d] = sum_j dS[i, j] en_t[j, d]   -> columns [256, 256 + Dp)
            const uint32_t id_v = idesc(128, Dp, 0, 1);
            for (int ks = 0; ks < 8; ++ks)
                umma_bf16(tmem + 256, umma_desc_sw128(dz0 + (ks >> 2) * BLOCK_BYTES + (ks & 3) * 32),
                          desc_mn(et0 + ks * 2048, BLOCK_BYTES, p.variant), id_v, (uint32_t)(ks > 0));
            // dq_t[j, d] = sum_i dS[i, j] en_v[i, d]   -> columns [0, Dp)
            const uint32_t id_t = idesc(128, Dp, 1, 1);
            for (int ks = 0; ks < 8; ++ks)
                umma_bf16(tmem, desc_mn(dz0 + ks * 2048, BLOCK_BYTES, p.variant), desc_mn(e0 + ks * 2048, BLOCK_BYTES, p.variant), id_t,
                          (uint32_t)(ks > 0));
            umma_commit(sm.bar_mma);
        }
        mbar_wait_sleepy(sm.bar_mma, mma_phase, 32);
        mma_phase ^= 1;
        tc_fence_after();
        const int half = Dp / 2;
        for (int mod = 0; mod < 2; ++mod) {
            const uint32_t col0 = (mod == 0 ? 256u : 0u) + (uint32_t)(h * half);
            const float* enr = p.en + ((int64_t)mod * N + (n < N ? n : 0)) * D;
            float dot = 0.f;
#pragma unroll 1
            for (int jj = 0; jj < half / 32; ++jj) {
                float v[32];
                tmem_ld32(tmem + lanes + col0 + (uint32_t)(jj * 32), v);
                const int d0 = h * half + jj * 32;
                if (n < N && d0 < D) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 e4 = *reinterpret_cast<const float4*>(enr + d0 + 4 * k);
                        dot = fmaf(v[4 * k], e4.x, dot); dot = fmaf(v[4 * k + 1], e4.y, dot);
                        dot = fmaf(v[4 * k + 2], e4.z, dot); dot = fmaf(v[4 * k + 3], e4.w, dot);
                    }
                }
            }
            __syncthreads();
            s_part[h * 128 + n] = dot;
            __syncthreads();
            dot = s_part[n] + s_part[128 + n];
            const float inv = n < N ? p.inv_e[mod * N + n] : 0.f;
            float* out = p.d_ga + ((int64_t)mod * N + (n < N ? n : 0)) * D;
#pragma unroll 1
            for (int jj = 0; jj < half / 32; ++jj) {
                float v[32];
                tmem_ld32(tmem + lanes + col0 + (uint32_t)(jj * 32), v);
                const int d0 = h * half + jj * 32;
                if (n < N && d0 < D) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 e4 = *reinterpret_cast<const float4*>(enr + d0 + 4 * k);
                        float4 o;
                        o.x = (v[4 * k] - dot * e4.x) * inv; o.y = (v[4 * k + 1] - dot * e4.y) * inv;
                        o.z = (v[4 * k + 2] - dot * e4.z) * inv; o.w = (v[4 * k + 3] - dot * e4.w) * inv;
                        *reinterpret_cast<float4*>(out + d0 + 4 * k) = o;
                    }
                }
            }
        }
    }
}

// after the second grid barrier: every CTA takes a share of the fixed-order partial reductions
__device__ void finish_phase(const FP& p, const Smem& sm) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int N = p.N, D = p.D, Dp = p.KC * 64, G = gridDim.x;
    if (p.want_grad && p.n_inst) {
        const int per_row = D / 4, total4 = 2 * N * per_row;
        const size_t tstride = (size_t)256 * Dp;
        for (int idx = blockIdx.x * F_THREADS + tid; idx < total4; idx += G * F_THREADS) {
            const int row = idx / per_row, d4 = idx % per_row;
            const int mod = row / N, nn = row % N;
            const float* src = p.part_inst + (size_t)(mod * 128 + nn) * Dp + d4 * 4;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            int t = 0;
            for (; t + 1 < p.T_inst; t += 2) {
                const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)t * tstride));
                const float4 y = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(t + 1) * tstride));
                a0.x += x.x; a0.y += x.y; a0.z += x.z; a0.w += x.w;
                a1.x += y.x; a1.y += y.y; a1.z += y.z; a1.w += y.w;
            }
            if (t < p.T_inst) {
                const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)t * tstride));
                a0.x += x.x; a0.y += x.y; a0.z += x.z; a0.w += x.w;
            }
            *reinterpret_cast<float4*>(p.d_inst + (size_t)idx * 4) = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
        }
    }
    if (p.want_grad && p.n_nce) {
        // d_nce[row] = normalise-backward( sum_tiles dq_part + dpos * key )   one warp per row
        for (int row = blockIdx.x * 8 + w; row < 2 * N; row += G * 8) {
            const int mod = row / N, nn = row % N;
            const float dp = __ldcg(p.dpos + row);
            const float* key = p.key_n[mod] + (int64_t)nn * D;
            const float* qr = p.qn + (int64_t)row * D;
            float4 g[2];
            float dot = 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int d4 = lane + 32 * u;
                g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (d4 * 4 < D) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float* src = p.part_nce + ((size_t)mod * p.T_k * 128 + nn) * Dp + d4 * 4;
                    for (int t = 0; t < p.T_k; ++t) {
                        const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)t * 128 * Dp));
                        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
                    }
                    const float4 k4 = *reinterpret_cast<const float4*>(key + d4 * 4);
                    const float4 q4 = *reinterpret_cast<const float4*>(qr + d4 * 4);
                    a.x = fmaf(dp, k4.x, a.x); a.y = fmaf(dp, k4.y, a.y); a.z = fmaf(dp, k4.z, a.z); a.w = fmaf(dp, k4.w, a.w);
                    dot = fmaf(a.x, q4.x, dot); dot = fmaf(a.y, q4.y, dot); dot = fmaf(a.z, q4.z, dot); dot = fmaf(a.w, q4.w, dot);
                    g[u] = a;
                }
            }
            dot = warp_sum(dot);
            const float inv = p.inv_q[row];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int d4 = lane + 32 * u;
                if (d4 * 4 < D) {
                    const float4 q4 = *reinterpret_cast<const float4*>(qr + d4 * 4);
                    float4 o;
                    o.x = (g[u].x - dot * q4.x) * inv; o.y = (g[u].y - dot * q4.y) * inv;
                    o.z = (g[u].z - dot * q4.z) * inv; o.w = (g[u].w - dot * q4.w) * inv;
                    *reinterpret_cast<float4*>(p.d_nce + (int64_t)row * D + d4 * 4) = o;
                }
            }
        }
    }
    if (p.reduce_losses && blockIdx.x == 0) {
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

__global__ void __launch_bounds__(F_THREADS, 1) fused_loss_kernel(const FP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    Smem sm;
    sm.E = smem + F_OFF_E; sm.DZ = smem + F_OFF_DZ; sm.WB = smem + F_OFF_WB;
    uint8_t* misc = smem + F_OFF_MISC;
    sm.inv = reinterpret_cast<float*>(misc);
    sm.valid = reinterpret_cast<uint32_t*>(misc + 512);
    sm.bar_load = reinterpret_cast<uint64_t*>(misc + 528);
    sm.bar_mma = reinterpret_cast<uint64_t*>(misc + 536);
    sm.tmem_slot = reinterpret_cast<uint32_t*>(misc + 544);
    sm.red32 = reinterpret_cast<float*>(misc + 576);
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        mbar_init(sm.bar_load, 1);
        mbar_init(sm.bar_mma, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(sm.tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sm.tmem_slot;

    const int b = blockIdx.x;
    if (b < p.n_inst) tile_program<true>(p, sm, tmem, 0, b);
    else if (b < p.n_inst + p.n_nce) tile_program<false>(p, sm, tmem, (b - p.n_inst) / p.T_k, (b - p.n_inst) % p.T_k);
    else align_program(p, sm, tmem);

    grid_barrier(p.bar + 1, gridDim.x);
    finish_phase(p, sm);

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// prologue: one warp per padded row (2 x 128).  fp32 artefacts as in loss_f32.cu's prologue (same expressions) plus the three
// packed bf16 operand images (raw embeds, normalised embeds, normalised InfoNCE queries), zero padded, and the barrier reset.
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fused_prologue_kernel(const float* __restrict__ v_embed, const float* __restrict__ t_embed, const float* __restrict__ v_qraw,
                      const float* __restrict__ t_qraw, const float* __restrict__ v_key, const float* __restrict__ t_key,
                      int normalize_keys, float* __restrict__ v_key_n, float* __restrict__ t_key_n, float* __restrict__ E2,
                      float* __restrict__ en, float* __restrict__ inv_e, float* __restrict__ qn, float* __restrict__ inv_q,
                      float* __restrict__ pos, uint8_t* __restrict__ Ep, uint8_t* __restrict__ ENp, uint8_t* __restrict__ QNp,
                      unsigned* __restrict__ bar, int N, int D, int KC) {
    const int lane = threadIdx.x & 31;
    const int prow = blockIdx.x * 8 + (threadIdx.x >> 5);         // padded row: modality * 128 + n
    if (blockIdx.x == 0 && threadIdx.x < 2) bar[threadIdx.x] = 0u;
    if (prow >= 256) return;
    const int mod = prow >> 7, n = prow & 127;
    const int k0 = lane * 8;
    const bool in_pad = k0 < KC * 64;
    float a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = b[i] = c[i] = 0.f;
    const bool live = n < N && k0 < D;
    const int row = mod * N + n;
    if (live) {
        const float* e = (mod ? t_embed : v_embed) + (int64_t)n * D + k0;
        const float* r = (mod ? t_qraw : v_qraw) + (int64_t)n * D + k0;
        const float* kin = (mod ? v_key : t_key) + (int64_t)n * D + k0;      // v queries pair with TEXT keys (head.py:160,166)
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = e[i]; b[i] = r[i]; c[i] = kin[i]; }
    }
    float se = 0.f, sr = 0.f, sk = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { se = fmaf(a[i], a[i], se); sr = fmaf(b[i], b[i], sr); sk = fmaf(c[i], c[i], sk); }
    se = warp_sum(se); sr = warp_sum(sr); sk = warp_sum(sk);
    const float ne = fmaxf(sqrtf(se), 1e-12f), nr = fmaxf(sqrtf(sr), 1e-12f);
    const float nk = normalize_keys ? fmaxf(sqrtf(sk), 1e-12f) : 1.0f;
    float an[8], qv[8], kv[8], dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        an[i] = __fdiv_rn(a[i], ne);
        qv[i] = __fdiv_rn(b[i], nr);
        kv[i] = normalize_keys ? __fdiv_rn(c[i], nk) : c[i];
        dot = fmaf(qv[i], kv[i], dot);
    }
    dot = warp_sum(dot);
    if (live) {
        float* kout = (mod ? v_key_n : t_key_n) + (int64_t)n * D + k0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            E2[(int64_t)row * D + k0 + i] = a[i];
            en[(int64_t)row * D + k0 + i] = an[i];
            qn[(int64_t)row * D + k0 + i] = qv[i];
            kout[i] = kv[i];
        }
    }
    if (n < N && lane == 0) { inv_e[row] = __fdiv_rn(1.0f, ne); inv_q[row] = __fdiv_rn(1.0f, nr); pos[row] = dot; }
    if (in_pad) {
        const int64_t off = packed_offset_bytes(prow, lane, KC);
        uint4 o;
        o.x = pack2(a[0], a[1]); o.y = pack2(a[2], a[3]); o.z = pack2(a[4], a[5]); o.w = pack2(a[6], a[7]);
        *reinterpret_cast<uint4*>(Ep + off) = o;
        o.x = pack2(an[0], an[1]); o.y = pack2(an[2], an[3]); o.z = pack2(an[4], an[5]); o.w = pack2(an[6], an[7]);
        *reinterpret_cast<uint4*>(ENp + off) = o;
        o.x = pack2(qv[0], qv[1]); o.y = pack2(qv[2], qv[3]); o.z = pack2(qv[4], qv[5]); o.w = pack2(qv[6], qv[7]);
        *reinterpret_cast<uint4*>(QNp + off) = o;
    }
}

struct Scratch {
    uint8_t *Ep, *ENp, *QNp;
    float4 *st_inst, *st_nce;
    float *part_inst, *part_nce;
    unsigned* bar;
    int64_t bytes;
};

Scratch carve_scratch(uint8_t* base, int N, int D, int K, int C) {
    (void)N;
    const int KC = (D + 127) / 128 * 2, Dp = KC * 64;
    const int T_inst = (C + F_TILE - 1) / F_TILE, T_k = (K + F_TILE - 1) / F_TILE;
    Scratch s;
    uint8_t* p = base;
    auto take = [&](int64_t bytes) { uint8_t* r = p; p += (bytes + 1023) / 1024 * 1024; return r; };
    const int64_t img = (int64_t)2 * KC * BLOCK_BYTES;
    s.Ep = take(img); s.ENp = take(img); s.QNp = take(img);
    s.st_inst = reinterpret_cast<float4*>(take((int64_t)256 * T_inst * 16));
    s.st_nce = reinterpret_cast<float4*>(take((int64_t)256 * T_k * 16));
    s.part_inst = reinterpret_cast<float*>(take((int64_t)T_inst * 256 * Dp * 4));
    s.part_nce = reinterpret_cast<float*>(take((int64_t)2 * T_k * 128 * Dp * 4));
    s.bar = reinterpret_cast<unsigned*>(take(256));
    s.bytes = p - base;
    return s;
}

}  // namespace

bool fused_loss_supported(int N, int D, int K, int C, int sm_count) {
    if (N < 1 || N > 128 || D < 64 || D > 256 || (D % 64) != 0) return false;
    const int ctas = (C + F_TILE - 1) / F_TILE + 2 * ((K + F_TILE - 1) / F_TILE) + 1;
    return ctas <= sm_count;
}

int64_t fused_loss_scratch_bytes(int N, int D, int K, int C) { return carve_scratch(nullptr, N, D, K, C).bytes + 1024; }

int fused_loss_prologue(const FusedLossArgs& a, cudaStream_t st) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a.scratch) + 1023) & ~uintptr_t(1023));
    const Scratch s = carve_scratch(base, a.N, a.D, a.K, a.C);
    const int KC = (a.D + 127) / 128 * 2;
    fused_prologue_kernel<<<32, 256, 0, st>>>(a.v_embed, a.t_embed, a.v_qraw, a.t_qraw, a.v_key, a.t_key, a.normalize_keys,
                                              a.v_key_n, a.t_key_n, a.E2, a.en, a.inv_e, a.qn, a.inv_q, a.pos, s.Ep, s.ENp,
                                              s.QNp, s.bar, a.N, a.D, KC);
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
    p.T_inst = (a.C + F_TILE - 1) / F_TILE; p.T_k = (a.K + F_TILE - 1) / F_TILE;
    p.roles = a.roles;
    p.n_inst = (a.roles & 1) ? p.T_inst : 0;
    p.n_nce = (a.roles & 2) ? 2 * p.T_k : 0;
    p.n_ga = (a.roles & 4) ? 1 : 0;
    p.want_grad = a.d_inst != nullptr;
    p.reduce_losses = a.reduce_losses;
    const char* var = getenv("TRB_FUSED_VARIANT");
    p.variant = var ? atoi(var) : 0;
    p.T = a.T; p.eps = a.eps; p.alpha = a.alpha; p.beta = a.beta; p.sp = a.sp; p.sn = a.sn;
    p.W = a.projection;
    p.queue[0] = a.t_queue; p.queue[1] = a.v_queue;
    p.key_n[0] = a.t_key_n; p.key_n[1] = a.v_key_n;
    p.labels = a.labels; p.id_queue = a.id_queue;
    p.Ep = s.Ep; p.ENp = s.ENp; p.QNp = s.QNp;
    p.en = a.en; p.qn = a.qn; p.inv_e = a.inv_e; p.inv_q = a.inv_q; p.pos = a.pos;
    p.st_inst = s.st_inst; p.st_nce = s.st_nce; p.part_inst = s.part_inst; p.part_nce = s.part_nce;
    p.dpos = a.dpos; p.rows_inst = a.rows_inst; p.rows_nce = a.rows_nce; p.rows_ga = a.rows_ga;
    p.losses = a.losses; p.d_inst = a.d_inst; p.d_nce = a.d_nce; p.d_ga = a.d_ga; p.d_proj = a.d_proj;
    p.bar = s.bar;
    const int grid = p.n_inst + p.n_nce + p.n_ga;
    if (grid == 0) return 0;

    static bool attr = false;
    if (!attr) {
        TRB_CUDA_OK(cudaFuncSetAttribute(fused_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
        attr = true;
    }
    // all CTAs meet at two grid barriers: the launch must be co-resident (cooperative), one CTA per SM
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(F_THREADS);
    cfg.dynamicSmemBytes = F_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = getenv("TRB_FUSED_NO_COOP") ? 0 : 1;
    TRB_CUDA_OK(cudaLaunchKernelEx(&cfg, fused_loss_kernel, p));
    return 0;
}
