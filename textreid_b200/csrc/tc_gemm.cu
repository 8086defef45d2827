// Generic bf16 tcgen05 GEMM for the MoCo loss path (precision = 1):
//     C[m, n] = nscale[n] * sum_k A(m, k) * B(k, n)          fp32 accumulate in TMEM, fp32 output
// with A and B given as strided fp32 matrices.  Two launches per operand-free call:
//   pack_strided_kernel : strided fp32 -> bf16 in the packed, pre-swizzled tile-major layout of tc_common.cuh
//                         (optionally scaled per k or per row; the contraction dimension is zero-padded to 64)
//   tc_gemm_kernel      : persistent CTAs over (m-tile, n-tile, k-split) items; producer warp streams 128x64 A blocks and
//                         256x64 B blocks with cp.async.bulk, one thread issues tcgen05.mma (M=128, N=256, K=16), eight
//                         epilogue warps drain the double-buffered TMEM accumulator to global memory.
// Split-K items write partial tiles (C + split * c_split) that the caller reduces in a fixed order.
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace {

using namespace tc;

constexpr int G_TILE_M = 128, G_TILE_N = 256, G_UMMA_K = 16;
constexpr int G_STAGE_BYTES = 3 * BLOCK_BYTES;      // A 128x64 + B 256x64 bf16 = 48 KiB
constexpr int G_STAGES = 4;
constexpr int G_THREADS = 384;                      // 4 control warps + 8 epilogue warps
constexpr int G_EPI_WARP0 = 4, G_EPI_WARPS = 8;

struct GemmTcParams {
    const uint8_t* A;        // packed [Mp, Kp]
    const uint8_t* B;        // packed [Np, Kp]
    float* C;
    int64_t ldc, c_split;
    int M, N, kchunks, ksplit;
    const float* nscale;
    int m_tiles, n_tiles, items;
};

struct Item { int mt, nt, ks, kc_lo, kc_hi; };

__device__ __forceinline__ Item item_of(const GemmTcParams& p, int it) {
    Item x;
    x.ks = it % p.ksplit;
    const int t = it / p.ksplit;
    x.nt = t % p.n_tiles;
    x.mt = t / p.n_tiles;
    x.kc_lo = (int)((int64_t)p.kchunks * x.ks / p.ksplit);
    x.kc_hi = (int)((int64_t)p.kchunks * (x.ks + 1) / p.ksplit);
    return x;
}

__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_kernel(const GemmTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)G_STAGES * G_STAGE_BYTES);
    uint64_t* s_full = bars;                     // [G_STAGES]
    uint64_t* s_empty = bars + G_STAGES;         // [G_STAGES]
    uint64_t* t_full = bars + 2 * G_STAGES;      // [2]
    uint64_t* t_empty = bars + 2 * G_STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < G_STAGES; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, G_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int KC = p.kchunks;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
                const Item x = item_of(p, it);
                for (int kc = x.kc_lo; kc < x.kc_hi; ++kc) {
                    mbar_wait(s_empty + stage, phase ^ 1);
                    mbar_expect_tx(s_full + stage, G_STAGE_BYTES);
                    uint8_t* dst = smem + (size_t)stage * G_STAGE_BYTES;
                    bulk_g2s(dst, p.A + ((size_t)x.mt * KC + kc) * BLOCK_BYTES, BLOCK_BYTES, s_full + stage);
                    bulk_g2s(dst + BLOCK_BYTES, p.B + ((size_t)(2 * x.nt) * KC + kc) * BLOCK_BYTES, BLOCK_BYTES, s_full + stage);
                    bulk_g2s(dst + 2 * BLOCK_BYTES, p.B + ((size_t)(2 * x.nt + 1) * KC + kc) * BLOCK_BYTES, BLOCK_BYTES, s_full + stage);
                    if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(G_TILE_M, G_TILE_N);
            int stage = 0, tbuf = 0;
            uint32_t phase = 0, tphase = 0;
            const uint32_t base = smem_u32(smem);
            for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
                const Item x = item_of(p, it);
                mbar_wait(t_empty + tbuf, tphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)tbuf * G_TILE_N;
                for (int kc = x.kc_lo; kc < x.kc_hi; ++kc) {
                    mbar_wait(s_full + stage, phase);
                    tc_fence_after();
                    const uint32_t a0 = base + (uint32_t)stage * G_STAGE_BYTES, b0 = a0 + BLOCK_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / G_UMMA_K; ++kk)
                        umma_bf16(d_tmem, umma_desc_sw128(a0 + kk * G_UMMA_K * 2), umma_desc_sw128(b0 + kk * G_UMMA_K * 2), idesc,
                                  (uint32_t)((kc > x.kc_lo) | (kk != 0)));
                    umma_commit(s_empty + stage);
                    if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(t_full + tbuf);
                tbuf ^= 1;
                if (tbuf == 0) tphase ^= 1;
            }
        }
    } else if (warp >= G_EPI_WARP0) {
        const int quarter = warp & 3, half = (warp - G_EPI_WARP0) >> 2;
        int tbuf = 0;
        uint32_t tphase = 0;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item x = item_of(p, it);
            const int m = x.mt * G_TILE_M + quarter * 32 + lane;
            float* crow = p.C + (int64_t)x.ks * p.c_split + (int64_t)m * p.ldc;
            mbar_wait(t_full + tbuf, tphase);
            tc_fence_after();
            if (x.kc_hi > x.kc_lo) {
#pragma unroll 1
                for (int chunk = 0; chunk < 4; ++chunk) {
                    float v[32];
                    __syncwarp();
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tbuf * G_TILE_N + half * 128 + chunk * 32), v);
                    const int n0 = x.nt * G_TILE_N + half * 128 + chunk * 32;
                    if (m < p.M && n0 < p.N) {
                        if (p.nscale) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= (n0 + j < p.N) ? p.nscale[n0 + j] : 0.f;
                        }
                        float* dst = crow + n0;
                        if (n0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < p.N) dst[j] = v[j];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + tbuf);
            tbuf ^= 1;
            if (tbuf == 0) tphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// strided fp32 -> packed bf16.  One CTA converts a 64-row x 64-k tile through shared memory so that global reads are
// coalesced whichever of the two source strides is 1.
__global__ void __launch_bounds__(256)
pack_strided_kernel(const float* __restrict__ src, int64_t s_row, int64_t s_k, int rows, int kdim, int rows_padded,
                    int kchunks, const float* __restrict__ kscale, const float* __restrict__ rscale, uint8_t* __restrict__ packed) {
    __shared__ float tile[64][65];
    const int r0 = blockIdx.x * 64, kc = blockIdx.y, k0 = kc * 64;
    const int t = threadIdx.x;
    const bool k_contig = (s_k == 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = t + i * 256;
        const int r = k_contig ? (idx >> 6) : (idx & 63);
        const int k = k_contig ? (idx & 63) : (idx >> 6);
        float v = 0.f;
        if (r0 + r < rows && k0 + k < kdim) {
            v = src[(int64_t)(r0 + r) * s_row + (int64_t)(k0 + k) * s_k];
            if (kscale) v *= kscale[k0 + k];
            if (rscale) v *= rscale[r0 + r];
        }
        tile[r][k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int j = t + i * 256;
        const int r = j >> 3, c = j & 7;
        if (r0 + r >= rows_padded) continue;
        uint4 out;
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(&out);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(tile[r][c * 8 + e]);
        *reinterpret_cast<uint4*>(packed + packed_offset_bytes(r0 + r, (int64_t)kc * 8 + c, kchunks)) = out;
    }
}

}  // namespace

int64_t tc_gemm_packed_bytes(int rows, int kdim) {
    const int64_t rp = ((int64_t)rows + 255) / 256 * 256, kp = ((int64_t)kdim + 63) / 64 * 64;
    return rp * kp * 2;
}

int tc_pack_strided(const float* src, int64_t s_row, int64_t s_k, int rows, int kdim, const float* kscale, const float* rscale,
                    void* packed, cudaStream_t st) {
    const int rp = (rows + 255) / 256 * 256, kchunks = (kdim + 63) / 64;
    dim3 grid(rp / 64, kchunks);
    pack_strided_kernel<<<grid, 256, 0, st>>>(src, s_row, s_k, rows, kdim, rp, kchunks, kscale, rscale, (uint8_t*)packed);
    TRB_LAUNCH_OK();
    return 0;
}

int tc_gemm_launch(const void* A_packed, const void* B_packed, float* C, int64_t ldc, int64_t c_split, int M, int N, int kdim,
                   int ksplit, const float* nscale, cudaStream_t st) {
    GemmTcParams p;
    p.A = (const uint8_t*)A_packed; p.B = (const uint8_t*)B_packed; p.C = C; p.ldc = ldc; p.c_split = c_split;
    p.M = M; p.N = N; p.kchunks = (kdim + 63) / 64;
    p.ksplit = ksplit < 1 ? 1 : (ksplit > p.kchunks ? p.kchunks : ksplit);
    p.nscale = nscale;
    p.m_tiles = (M + G_TILE_M - 1) / G_TILE_M; p.n_tiles = (N + G_TILE_N - 1) / G_TILE_N;
    p.items = p.m_tiles * p.n_tiles * p.ksplit;
    if (p.items == 0) return 0;
    const int smem_bytes = G_STAGES * G_STAGE_BYTES + (2 * G_STAGES + 4) * 8 + 16 + 1024;
    static TrbDeviceOnce attr;
    if (trb_first_on_device(attr))
        TRB_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(p.items < sms ? p.items : sms);
    tc_gemm_kernel<<<grid, G_THREADS, smem_bytes, st>>>(p);
    TRB_LAUNCH_OK();
    return 0;
}
