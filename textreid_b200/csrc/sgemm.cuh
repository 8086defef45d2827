// Generic strided fp32 GEMM on the FFMA pipe (parity path and small helper contractions).
// Each output element is accumulated as acc = fmaf(a, b, acc) with k strictly ascending from
// acc = 0 (within a split), the same order as the retrieval stream / threshold kernels.
#pragma once
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------
// generic strided SGEMM: C[z][m,n] = nscale[n] * sum_k kscale[k] * A(m,k) * B(k,n)
// grid.z = split-K partials (each written to C + z*c_split), reduced afterwards
// ---------------------------------------------------------------------------
constexpr int GM = 64, GN = 64, GK = 16;

struct GemmArgs {
    const float* A; int64_t a_sm, a_sk;
    const float* B; int64_t b_sk, b_sn;
    float* C; int64_t c_sm, c_split;
    int M, N, K;
    const float* kscale;
    const float* nscale;
    int splitk;
};

static __global__ void __launch_bounds__(256) sgemm_generic_kernel(GemmArgs g) {
    __shared__ float As[GK][GM + 4];
    __shared__ float Bs[GK][GN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int kchunks = (g.K + GK - 1) / GK;
    const int c_lo = (int)((int64_t)kchunks * blockIdx.z / g.splitk);
    const int c_hi = (int)((int64_t)kchunks * (blockIdx.z + 1) / g.splitk);
    float acc[4][4] = {};
    for (int ch = c_lo; ch < c_hi; ++ch) {
        const int k0 = ch * GK;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int idx = tid + p * 256;
            int mm, kk;
            if (g.a_sk == 1) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
            float v = 0.f;
            if (m0 + mm < g.M && k0 + kk < g.K) {
                v = g.A[(int64_t)(m0 + mm) * g.a_sm + (int64_t)(k0 + kk) * g.a_sk];
                if (g.kscale) v *= g.kscale[k0 + kk];
            }
            As[kk][mm] = v;
            int nn;
            if (g.b_sn == 1) { nn = idx & 63; kk = idx >> 6; } else { kk = idx & 15; nn = idx >> 4; }
            v = 0.f;
            if (n0 + nn < g.N && k0 + kk < g.K) v = g.B[(int64_t)(k0 + kk) * g.b_sk + (int64_t)(n0 + nn) * g.b_sn];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = g.C + (int64_t)blockIdx.z * g.c_split;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.nscale) v *= g.nscale[n];
            C[(int64_t)m * g.c_sm + n] = v;
        }
    }
}

static inline int launch_gemm(const GemmArgs& g, cudaStream_t st) {
    dim3 grid((g.N + GN - 1) / GN, (g.M + GM - 1) / GM, g.splitk);
    sgemm_generic_kernel<<<grid, 256, 0, st>>>(g);
    TRB_LAUNCH_OK();
    return 0;
}


}  // namespace
