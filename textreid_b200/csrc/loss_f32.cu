// fp32 (FFMA) MoCo loss dict + gradients: the parity path (1e-5) of the north star.
//
// Math (SURVEY.md section 8a', verified against the reference's autograd):
//   mask[k]  = any_n(id_queue[k] == y[n])                                   head.py:148-157
//   InfoNCE  z = [<q,key> | q@queue (masked -> -inf)] / T, CE target 0       head.py:160-170, losses.py:206-217
//   instance z = e @ (W/||W||_col), label-smoothed CE                        losses.py:42-62, 6-39
//   align    S = q_v q_t^T, log(1+exp(.)) on same-id / other pairs           losses.py:102-128
// Forward and backward are one stream-ordered launch sequence: logits are materialised in the
// caller's workspace (a few MB), turned into their gradients in place, and contracted back.
// The bf16 tcgen05 path (precision = 1) reuses this sequence with tensor-core GEMMs (tc_gemm.cu) or, when the shape fits, replaces
// it by the fused cooperative kernel of loss_fused.cu, which keeps logits and their gradients on chip.
#include "common.cuh"
#include "sgemm.cuh"
#include "tc_gemm.cuh"
#include "loss_fused.cuh"
#include <stdlib.h>

namespace {

__global__ void reduce_partials_kernel(float* __restrict__ out, const float* __restrict__ part, int nsplit, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < nsplit; ++z) s += part[(int64_t)z * n + i];
    out[i] = s;
}

// ---------------------------------------------------------------------------
// prologue: normalised rows, positive logits, stacked embeds, queue column mask, column norms
// ---------------------------------------------------------------------------
// rows [0,N): v ; [N,2N): t.  One warp per row.
__global__ void __launch_bounds__(256)
prologue_rows_kernel(const float* __restrict__ v_embed, const float* __restrict__ t_embed, const float* __restrict__ v_qraw,
                     const float* __restrict__ t_qraw, const float* __restrict__ v_key, const float* __restrict__ t_key,
                     int normalize_keys, float* __restrict__ v_key_n, float* __restrict__ t_key_n, float* __restrict__ E2,
                     float* __restrict__ en, float* __restrict__ inv_e, float* __restrict__ qn, float* __restrict__ inv_q,
                     float* __restrict__ pos, int N, int D) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= 2 * N) return;
    const int mod = row / N, n = row % N;
    const float* e = (mod ? t_embed : v_embed) + (int64_t)n * D;
    const float* r = (mod ? t_qraw : v_qraw) + (int64_t)n * D;
    // v queries pair with TEXT keys, t queries with IMAGE keys (head.py:160,166)
    const float* kin = (mod ? v_key : t_key) + (int64_t)n * D;
    float* kout = (mod ? v_key_n : t_key_n) + (int64_t)n * D;
    float se = 0.f, sr = 0.f, sk = 0.f;
    for (int k = lane; k < D; k += 32) {
        const float a = e[k], b = r[k], c = kin[k];
        se = fmaf(a, a, se); sr = fmaf(b, b, sr); sk = fmaf(c, c, sk);
    }
    se = warp_sum(se); sr = warp_sum(sr); sk = warp_sum(sk);
    const float ne = fmaxf(sqrtf(se), 1e-12f), nr = fmaxf(sqrtf(sr), 1e-12f);
    const float nk = normalize_keys ? fmaxf(sqrtf(sk), 1e-12f) : 1.0f;
    float dot = 0.f;
    for (int k = lane; k < D; k += 32) {
        const float a = e[k];
        const float qv = __fdiv_rn(r[k], nr);
        const float kv = normalize_keys ? __fdiv_rn(kin[k], nk) : kin[k];
        E2[(int64_t)row * D + k] = a;
        en[(int64_t)row * D + k] = __fdiv_rn(a, ne);
        qn[(int64_t)row * D + k] = qv;
        kout[k] = kv;
        dot = fmaf(qv, kv, dot);
    }
    dot = warp_sum(dot);
    if (lane == 0) { inv_e[row] = __fdiv_rn(1.0f, ne); inv_q[row] = __fdiv_rn(1.0f, nr); pos[row] = dot; }
}

__global__ void __launch_bounds__(256)
queue_mask_kernel(const int64_t* __restrict__ id_queue, const int64_t* __restrict__ labels, uint8_t* __restrict__ mask, int N, int K) {
    extern __shared__ int64_t lab[];
    for (int i = threadIdx.x; i < N; i += blockDim.x) lab[i] = labels[i];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int64_t id = id_queue[k];
    bool hit = false;
    for (int n = 0; n < N; ++n) hit |= (lab[n] == id);
    mask[k] = hit ? 1 : 0;
}

// 32 columns x 8 row-groups per CTA: coalesced along the class dimension, 8-way split of the D-long column walk
__global__ void __launch_bounds__(256)
column_inv_norm_kernel(const float* __restrict__ W, float* __restrict__ inv_c, int D, int C) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float ss = 0.f;
    if (c < C)
        for (int d = ty; d < D; d += 8) { const float w = W[(int64_t)d * C + c]; ss = fmaf(w, w, ss); }
    part[ty][tx] = ss;
    __syncthreads();
    if (ty == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][tx];
        inv_c[c] = __fdiv_rn(1.0f, fmaxf(sqrtf(t), 1e-12f));
    }
}

// ---------------------------------------------------------------------------
// row kernels: logits -> loss contribution and logits gradient in place
// ---------------------------------------------------------------------------
// InfoNCE (losses.py:206-217).  S row [K] holds q@queue; column 0 of the reference's logits is pos.
__global__ void __launch_bounds__(256)
nce_rows_kernel(float* __restrict__ S, const float* __restrict__ pos, const uint8_t* __restrict__ mask, float T, int N, int K,
                float* __restrict__ loss_row, float* __restrict__ dpos, int want_grad) {
    __shared__ float red[32];
    const int row = blockIdx.x;
    float* s = S + (int64_t)row * K;
    const float z0 = __fdiv_rn(pos[row], T);
    float mx = z0;
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (!mask[k]) mx = fmaxf(mx, __fdiv_rn(s[k], T));
    mx = block_max(mx, red);
    float se = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (!mask[k]) se += expf(__fdiv_rn(s[k], T) - mx);
    se = block_sum(se, red);
    se += expf(z0 - mx);
    const float lse = mx + logf(se);
    if (threadIdx.x == 0) loss_row[row] = lse - z0;
    if (!want_grad) return;
    const float gscale = 1.0f / ((float)N * T);
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        s[k] = mask[k] ? 0.f : expf(__fdiv_rn(s[k], T) - lse) * gscale;
    if (threadIdx.x == 0) dpos[row] = (expf(z0 - lse) - 1.0f) * gscale;
}

// instance loss with label smoothing (losses.py:26-39, 53-60).  Z row [C].  One sweep gathers (max, sum exp, sum z)
// with the online-softmax recurrence, a second sweep writes the logit gradient in place.
__global__ void __launch_bounds__(256)
instance_rows_kernel(float* __restrict__ Z, const int64_t* __restrict__ labels, float eps, int N, int C,
                     float* __restrict__ loss_row, int want_grad) {
    __shared__ float red[32];
    const int row = blockIdx.x;
    float* z = Z + (int64_t)row * C;
    // a label outside [0, C) makes the reference raise (scatter_ in CrossEntropyLabelSmooth, losses.py:33); here the row's loss
    // becomes NaN (and no class is treated as the target) instead of an out-of-bounds read
    const int64_t lab = labels[row % N];
    const bool bad_label = lab < 0 || lab >= (int64_t)C;
    const int y = bad_label ? -1 : (int)lab;
    float m = -CUDART_INF_F, se = 0.f, sz = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = z[c];
        sz += v;
        if (v > m) { se = se * expf(m - v) + 1.0f; m = v; }
        else se += expf(v - m);
    }
    const float mx = block_max(m, red);
    se = block_sum(se * expf(m - mx), red);       // threads with no element carry m = -inf, se = 0 -> contribute 0
    sz = block_sum(sz, red);
    const float lse = mx + logf(se);
    const float zy = bad_label ? CUDART_NAN_F : z[y];
    __syncthreads();
    if (threadIdx.x == 0) loss_row[row] = lse - (1.0f - eps) * zy - (eps / (float)C) * sz;
    if (!want_grad) return;
    const float invN = 1.0f / (float)N, uni = eps / (float)C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float t = uni + ((c == y) ? (1.0f - eps) : 0.f);
        z[c] = (expf(z[c] - lse) - t) * invN;
    }
}

// global align (losses.py:114-127): S [N,N] -> per-row loss sums, dS in place
__global__ void __launch_bounds__(128)
align_rows_kernel(float* __restrict__ S, const int64_t* __restrict__ labels, float alpha, float beta, float sp, float sn,
                  int N, float* __restrict__ loss_row, int want_grad) {
    __shared__ float red[32];
    const int i = blockIdx.x;
    const int64_t yi = labels[i];
    const float two_over_n = 2.0f / (float)N;
    float acc = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float s = S[(int64_t)i * N + j];
        const bool same = labels[j] == yi;
        const float x = same ? -sp * (s - alpha) : sn * (s - beta);
        const float e = expf(x);
        acc += logf(1.0f + e);
        if (want_grad) S[(int64_t)i * N + j] = (same ? -sp : sn) * (e / (1.0f + e)) * two_over_n;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) loss_row[i] = acc;
}

// de = (dq + dpos*key - <.,q> q) * inv_norm : backward of x -> x/||x||   (one warp per row)
__global__ void __launch_bounds__(256)
normalize_backward_kernel(const float* __restrict__ dq, const float* __restrict__ dpos, const float* __restrict__ key_v,
                          const float* __restrict__ key_t, const float* __restrict__ qn, const float* __restrict__ inv_norm,
                          float* __restrict__ out, int N, int D) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= 2 * N) return;
    const float* g = dq + (int64_t)row * D;
    const float* q = qn + (int64_t)row * D;
    const float* key = nullptr;
    float dp = 0.f;
    if (dpos) { dp = dpos[row]; key = (row < N ? key_t : key_v) + (int64_t)(row % N) * D; }
    float dot = 0.f;
    for (int k = lane; k < D; k += 32) {
        const float gv = g[k] + (key ? dp * key[k] : 0.f);
        dot = fmaf(gv, q[k], dot);
    }
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
    for (int k = lane; k < D; k += 32) {
        const float gv = g[k] + (key ? dp * key[k] : 0.f);
        out[(int64_t)row * D + k] = (gv - dot * q[k]) * inv;
    }
}

// dW = (dWhat - <dWhat, What>_col What) / ||W||_col, What = W * inv_c   (in place; 32 columns x 8 row-groups per CTA)
__global__ void __launch_bounds__(256)
projection_backward_kernel(float* __restrict__ dW, const float* __restrict__ W, const float* __restrict__ inv_c, int D, int C) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const float ic = c < C ? inv_c[c] : 0.f;
    float dot = 0.f;
    if (c < C)
        for (int d = ty; d < D; d += 8) dot = fmaf(dW[(int64_t)d * C + c], W[(int64_t)d * C + c] * ic, dot);
    part[ty][tx] = dot;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    if (c < C)
        for (int d = ty; d < D; d += 8) {
            const int64_t o = (int64_t)d * C + c;
            dW[o] = (dW[o] - t * (W[o] * ic)) * ic;
        }
}

// losses[0] = sum(inst rows)/N ; losses[1] = sum(nce rows)/N ; losses[2] = sum(align rows)*2/N
__global__ void __launch_bounds__(256)
loss_reduce_kernel(const float* __restrict__ inst_rows, const float* __restrict__ nce_rows, const float* __restrict__ ga_rows,
                   int N, float* __restrict__ losses) {
    __shared__ float red[32];
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) { a += inst_rows[i]; b += nce_rows[i]; }
    for (int i = threadIdx.x; i < N; i += blockDim.x) c += ga_rows[i];
    a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red);
    if (threadIdx.x == 0) {
        losses[0] = a / (float)N;
        losses[1] = b / (float)N;
        losses[2] = c * 2.0f / (float)N;
    }
}

constexpr int SPLIT_NCE = 8, SPLIT_INST = 32;

struct Workspace {
    float *E2, *en, *qn, *inv_e, *inv_q, *pos, *dpos, *S_nce, *Z, *inv_c, *S_ga, *dq_nce, *dq_ga, *part, *rows_inst,
        *rows_nce, *rows_ga;
    uint8_t* mask;
    uint8_t *pkA, *pkB;       // packed bf16 operand scratch of the tensor-core path (instance branch)
    uint8_t *pkA_nce, *pkB_nce, *pkA_ga, *pkB_ga, *pkA_dw, *pkB_dw;
    float* part_nce;          // split-K partials of the InfoNCE branch
    uint8_t* fused;           // scratch of the fused cooperative kernel (loss_fused.cu); null when the shape does not fit it
    int64_t bytes;
};

// SM count of the current device (148 when there is none, e.g. a size query in a CPU-only build container)
static int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        return 148;
    }
    return sms;
}

// split-K factor actually used: the tensor-core kernel cannot split finer than one 64-wide k-chunk per part
static inline int eff_split(int split, int K, bool use_tc) {
    const int chunks = use_tc ? (K + 63) / 64 : (K + GK - 1) / GK;
    return split < chunks ? split : (chunks < 1 ? 1 : chunks);
}

// One contraction, on the FFMA pipe (precision 0) or on the tensor cores (precision 1: both operands are first
// rounded to bf16 into the packed tile-major layout, then tcgen05.mma accumulates in fp32).
struct Scratch { uint8_t* pkA; uint8_t* pkB; };
int run_gemm(const GemmArgs& g, bool use_tc, const Scratch& sc, cudaStream_t st) {
    if (!use_tc) return launch_gemm(g, st);
    int rc;
    if ((rc = tc_pack_strided(g.A, g.a_sm, g.a_sk, g.M, g.K, g.kscale, nullptr, sc.pkA, st))) return rc;
    if ((rc = tc_pack_strided(g.B, g.b_sn, g.b_sk, g.N, g.K, nullptr, nullptr, sc.pkB, st))) return rc;
    return tc_gemm_launch(sc.pkA, sc.pkB, g.C, g.c_sm, g.c_split, g.M, g.N, g.K, g.splitk, g.nscale, st);
}

// The three loss branches (instance / InfoNCE / global-align) are independent between the prologue and the final
// reduction; they run on the caller's stream plus two helper streams, forked and joined with events, so that a
// captured CUDA graph becomes a 3-wide DAG instead of a chain of ~45 dependent small launches.  The helper streams and
// events are process-level, created on first use (per device), and carry no data between calls.
struct Fork { cudaStream_t s1, s2, s3; cudaEvent_t fork, fork3, j1, j2, j3; bool ok; };
static Fork* helper_streams() {
    static Fork pool[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    Fork& f = pool[dev];
    if (!f.ok) {
        if (cudaStreamCreateWithFlags(&f.s1, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&f.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&f.s3, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f.fork3, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f.j3, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f.j1, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f.j2, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        f.ok = true;
    }
    return &f;
}

Workspace carve(void* base, int N, int D, int K, int C, bool use_tc) {
    Workspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](int64_t nfloats) {
        float* r = reinterpret_cast<float*>(p);
        p += ((nfloats * 4 + 255) / 256) * 256;
        return r;
    };
    const int64_t ND2 = 2LL * N * D;
    w.E2 = take(ND2); w.en = take(ND2); w.qn = take(ND2);
    w.inv_e = take(2 * N); w.inv_q = take(2 * N); w.pos = take(2 * N); w.dpos = take(2 * N);
    w.S_nce = take(2LL * N * K);
    w.Z = take(2LL * N * C);
    w.inv_c = take(C);
    w.S_ga = take((int64_t)N * N);
    w.dq_nce = take(ND2); w.dq_ga = take(ND2);
    const int64_t part = (int64_t)(SPLIT_INST > SPLIT_NCE ? SPLIT_INST : SPLIT_NCE) * ND2;
    w.part = take(part);
    w.rows_inst = take(2 * N); w.rows_nce = take(2 * N); w.rows_ga = take(N);
    w.mask = reinterpret_cast<uint8_t*>(take((K + 3) / 4));
    w.pkA = w.pkB = nullptr;
    if (use_tc) {
        // largest packed operands: A = dZ [2N x C], B = projection^T [C x D] / dZ^T [C x 2N] / queues [K x D]
        int64_t a = tc_gemm_packed_bytes(2 * N, C), b = tc_gemm_packed_bytes(C, D);
        const int64_t cand[] = {tc_gemm_packed_bytes(2 * N, D), tc_gemm_packed_bytes(2 * N, K), tc_gemm_packed_bytes(D, 2 * N)};
        for (int64_t x : cand) a = x > a ? x : a;
        const int64_t candb[] = {tc_gemm_packed_bytes(C, 2 * N), tc_gemm_packed_bytes(K, D), tc_gemm_packed_bytes(D, K),
                                 tc_gemm_packed_bytes(D, C), tc_gemm_packed_bytes(D, N), tc_gemm_packed_bytes(N, D)};
        for (int64_t x : candb) b = x > b ? x : b;
        w.pkA = reinterpret_cast<uint8_t*>(take(a / 4 + 64));
        w.pkB = reinterpret_cast<uint8_t*>(take(b / 4 + 64));
        const int64_t an = tc_gemm_packed_bytes(N, K) > tc_gemm_packed_bytes(N, D) ? tc_gemm_packed_bytes(N, K) : tc_gemm_packed_bytes(N, D);
        const int64_t bn = tc_gemm_packed_bytes(K, D) > tc_gemm_packed_bytes(D, K) ? tc_gemm_packed_bytes(K, D) : tc_gemm_packed_bytes(D, K);
        w.pkA_nce = reinterpret_cast<uint8_t*>(take(an / 4 + 64));
        w.pkB_nce = reinterpret_cast<uint8_t*>(take(bn / 4 + 64));
        const int64_t ag = tc_gemm_packed_bytes(N, N) > tc_gemm_packed_bytes(N, D) ? tc_gemm_packed_bytes(N, N) : tc_gemm_packed_bytes(N, D);
        const int64_t bg = tc_gemm_packed_bytes(N, D) > tc_gemm_packed_bytes(D, N) ? tc_gemm_packed_bytes(N, D) : tc_gemm_packed_bytes(D, N);
        w.pkA_ga = reinterpret_cast<uint8_t*>(take(ag / 4 + 64));
        w.pkB_ga = reinterpret_cast<uint8_t*>(take(bg / 4 + 64));
        w.pkA_dw = reinterpret_cast<uint8_t*>(take(tc_gemm_packed_bytes(D, 2 * N) / 4 + 64));
        w.pkB_dw = reinterpret_cast<uint8_t*>(take(tc_gemm_packed_bytes(C, 2 * N) / 4 + 64));
    } else {
        w.pkA_nce = w.pkB_nce = w.pkA_ga = w.pkB_ga = w.pkA_dw = w.pkB_dw = nullptr;
    }
    w.part_nce = take((int64_t)SPLIT_NCE * N * D);
    w.fused = nullptr;
    if (use_tc && (fused_loss_supported(N, D, K, C, sm_count()) || fused_windows_supported(N, D, K, C, sm_count())))
        w.fused = reinterpret_cast<uint8_t*>(take(fused_loss_scratch_bytes(N, D, K, C) / 4 + 64));
    w.bytes = p - static_cast<char*>(base);
    return w;
}

}  // namespace

int64_t trb_moco_loss_workspace_bytes_f32(const trb_moco_shape* s) {
    return carve(nullptr, s->N, s->D, s->K, s->C, false).bytes;
}
int64_t trb_moco_loss_workspace_bytes_tc(const trb_moco_shape* s) {
    return carve(nullptr, s->N, s->D, s->K, s->C, true).bytes;
}

// launches of one call (grads and d_projection requested); mirrors the sequence below
int trb_moco_loss_launches_impl(const trb_moco_shape* s, int precision) {
    const bool use_tc = precision == 1;
    if (use_tc && fused_loss_supported(s->N, s->D, s->K, s->C, sm_count())) {
        const char* e = getenv("TRB_FUSED_ROLES");
        if (e == nullptr || (atoi(e) & 7) == 7) return 2;
    }
    const int per_gemm = use_tc ? 3 : 1;
    if (use_tc && fused_windows_supported(s->N, s->D, s->K, s->C, sm_count())) {
        const char* e = getenv("TRB_FUSED_ROLES");
        if (e == nullptr || (atoi(e) & 3) == 3) {
            // shared prologue + the unfused global-align branch (3 contractions, pair losses, normalise backward) + the fused
            // prologue and one cooperative launch (the row windows are walked inside the kernel) + loss reduce
            return 1 + (3 * per_gemm + 2) + 2 + 1;
        }
    }
    // prologue, mask, column norms, 3 row kernels, 2 normalise-backward, projection backward, loss reduce, 3 partial reductions
    return 13 + 10 * per_gemm;
}

// `queue_ptr` != NULL: the step ends with _dequeue_and_enqueue (head.py:175); the queues are then written (after every read)
static int moco_loss_impl(const float* v_embed, const float* t_embed, const float* v_qraw, const float* t_qraw,
                          const float* v_key, const float* t_key, int normalize_keys, float* v_key_n, float* t_key_n,
                          const int64_t* labels, const float* v_queue, const float* t_queue, const int64_t* id_queue,
                          const float* projection, const trb_moco_shape* shape, const trb_moco_hparams* hp, float* losses,
                          float* d_inst, float* d_nce, float* d_ga, float* d_projection, void* workspace,
                          int64_t workspace_bytes, cudaStream_t st, bool use_tc, int64_t* queue_ptr) {
    const int N = shape->N, D = shape->D, K = shape->K, C = shape->C;
    Workspace w = carve(workspace, N, D, K, C, use_tc);
    if (workspace_bytes < w.bytes) {
        trb_set_error("moco_loss: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)w.bytes);
        return TRB_ERR_WORKSPACE;
    }
    const bool grads = d_inst != nullptr;
    const int rows = 2 * N;

    Fork* fk = helper_streams();
    if (fk == nullptr) { trb_set_error("moco_loss: could not create helper streams"); return TRB_ERR_INVALID; }
    cudaStream_t s_nce = fk->s1, s_ga = fk->s2;          // the instance branch stays on the caller's stream
    const Scratch sc_inst{w.pkA, w.pkB}, sc_nce{w.pkA_nce, w.pkB_nce}, sc_ga{w.pkA_ga, w.pkB_ga}, sc_dw{w.pkA_dw, w.pkB_dw};
    const int64_t ND = (int64_t)N * D;
    const int split_nce = eff_split(SPLIT_NCE, K, use_tc), split_inst = eff_split(SPLIT_INST, C, use_tc);
    int rc;

    // ---- fused path: when the shape fits, the prologue plus ONE cooperative tcgen05 kernel (loss_fused.cu) replace the
    //      branches named in `roles` (bit 0 instance, 1 InfoNCE, 2 align; TRB_FUSED_ROLES narrows it for debugging)
    int roles = 0;
    FusedLossArgs fa;
    memset(&fa, 0, sizeof(fa));
    // up to 256 rows everything is one cooperative launch; above that (up to 1024) the instance and InfoNCE branches are, and
    // the global-align branch, whose N x N similarity couples all 128-row windows, stays on the unfused tensor-core sequence
    const bool windows = w.fused != nullptr && !fused_loss_supported(N, D, K, C, sm_count());
    if (w.fused != nullptr) {
        roles = windows ? 3 : 7;
        if (const char* e = getenv("TRB_FUSED_ROLES")) roles = windows ? (((atoi(e) & 3) == 3) ? 3 : 0) : (atoi(e) & 7);
        // vector accesses of the fused kernels: 16-byte aligned inputs, 32-byte aligned embedding gradients
        const void* in16[] = {v_embed, t_embed, v_qraw, t_qraw, v_key, t_key, v_key_n, t_key_n};
        for (const void* q : in16)
            if (!trb_aligned16(q)) roles = 0;
        const void* out32[] = {d_inst, d_nce, d_ga};
        for (const void* q : out32)
            if (reinterpret_cast<uintptr_t>(q) & 31u) roles = 0;
        if (windows && (D % 8) != 0) roles = 0;
    }
    if (roles) {
        fa.N = N; fa.D = D; fa.K = K; fa.C = C;
        fa.T = hp->T; fa.eps = hp->epsilon; fa.alpha = hp->alpha; fa.beta = hp->beta; fa.sp = hp->scale_pos; fa.sn = hp->scale_neg;
        fa.v_embed = v_embed; fa.t_embed = t_embed; fa.v_qraw = v_qraw; fa.t_qraw = t_qraw; fa.v_key = v_key; fa.t_key = t_key;
        fa.normalize_keys = normalize_keys; fa.labels = labels; fa.id_queue = id_queue;
        fa.v_queue = v_queue; fa.t_queue = t_queue; fa.projection = projection;
        fa.v_key_n = v_key_n; fa.t_key_n = t_key_n; fa.E2 = w.E2; fa.en = w.en; fa.qn = w.qn; fa.inv_e = w.inv_e; fa.inv_q = w.inv_q;
        fa.pos = w.pos; fa.dpos = w.dpos; fa.rows_inst = w.rows_inst; fa.rows_nce = w.rows_nce; fa.rows_ga = w.rows_ga;
        fa.losses = losses; fa.d_inst = d_inst; fa.d_nce = d_nce; fa.d_ga = d_ga; fa.d_proj = d_projection;
        fa.scratch = w.fused; fa.roles = roles; fa.reduce_losses = roles == 7;
    }
    if (roles && !windows) {
        // the whole step fused: the enqueue rides in the cooperative kernel (its queue reads ended with the prologue's re-pack)
        const bool enq_inside = roles == 7 && queue_ptr != nullptr;
        fa.enq_v_queue = enq_inside ? const_cast<float*>(v_queue) : nullptr;
        fa.enq_t_queue = enq_inside ? const_cast<float*>(t_queue) : nullptr;
        fa.enq_ids = enq_inside ? const_cast<int64_t*>(id_queue) : nullptr;
        fa.enq_ptr = enq_inside ? queue_ptr : nullptr;
        if ((rc = fused_loss_prologue(fa, st))) return rc;
        if (roles == 7) {                                        // the whole step: prologue + one cooperative launch, no helper streams
            fa.after_prologue = 1;
            return fused_loss_launch(fa, st);
        }
    } else {
        // ---- shared prologue on the caller's stream
        prologue_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(v_embed, t_embed, v_qraw, t_qraw, v_key, t_key, normalize_keys,
                                                              v_key_n, t_key_n, w.E2, w.en, w.inv_e, w.qn, w.inv_q, w.pos, N, D);
        TRB_LAUNCH_OK();
    }
    TRB_CUDA_OK(cudaEventRecord(fk->fork, st));
    TRB_CUDA_OK(cudaStreamWaitEvent(s_nce, fk->fork, 0));
    TRB_CUDA_OK(cudaStreamWaitEvent(s_ga, fk->fork, 0));
    if (roles && !windows && (rc = fused_loss_launch(fa, st))) return rc;
    if (roles && windows) {
        // ---- batches above 128 rows on the caller's stream: ONE prologue (operand images of every 128-row window, queue images)
        //      and ONE cooperative launch that walks the windows inside the kernel with its W / queue tiles resident; an InfoNCE
        //      CTA takes its tile of both modalities in turn (86 instance + 2 x 32 InfoNCE tiles would not fit 148 SMs)
        fa.reduce_losses = 0;
        fa.roles = 3;
        if ((rc = fused_loss_prologue(fa, st))) return rc;
        fa.after_prologue = 1;
        if ((rc = fused_loss_launch(fa, st))) return rc;
        fa.after_prologue = 0;
        fa.roles = roles;
    }

    // ---- InfoNCE branch (helper stream 1): mask, logits of v queries x text queue and t queries x image queue
    //      (head.py:148-170), row-wise CE (losses.py:206-217), dq = dS @ queue^T + dpos * key, normalise backward
    if (!(roles & 2)) {
    queue_mask_kernel<<<(K + 255) / 256, 256, N * sizeof(int64_t), s_nce>>>(id_queue, labels, w.mask, N, K);
    TRB_LAUNCH_OK();
    for (int mod = 0; mod < 2; ++mod) {
        GemmArgs g{w.qn + (int64_t)mod * N * D, D, 1, mod ? v_queue : t_queue, K, 1,
                   w.S_nce + (int64_t)mod * N * K, K, 0, N, K, D, nullptr, nullptr, 1};
        if ((rc = run_gemm(g, use_tc, sc_nce, s_nce))) return rc;
    }
    nce_rows_kernel<<<rows, 256, 0, s_nce>>>(w.S_nce, w.pos, w.mask, hp->T, N, K, w.rows_nce, w.dpos, grads);
    TRB_LAUNCH_OK();
    if (grads) {
        for (int mod = 0; mod < 2; ++mod) {
            GemmArgs g{w.S_nce + (int64_t)mod * N * K, K, 1, mod ? v_queue : t_queue, 1, K,
                       w.part_nce, D, ND, N, D, K, nullptr, nullptr, split_nce};
            if ((rc = run_gemm(g, use_tc, sc_nce, s_nce))) return rc;
            reduce_partials_kernel<<<(unsigned)((ND + 255) / 256), 256, 0, s_nce>>>(w.dq_nce + mod * ND, w.part_nce, split_nce, ND);
            TRB_LAUNCH_OK();
        }
        normalize_backward_kernel<<<(rows + 7) / 8, 256, 0, s_nce>>>(w.dq_nce, w.dpos, v_key_n, t_key_n, w.qn, w.inv_q, d_nce, N, D);
        TRB_LAUNCH_OK();
    }
    }
    TRB_CUDA_OK(cudaEventRecord(fk->j1, s_nce));

    // ---- global-align branch (helper stream 2): S = q_v q_t^T (losses.py:114), pair losses, dq_v = dS q_t, dq_t = dS^T q_v
    if (!(roles & 4)) {
    {
        GemmArgs g{w.en, D, 1, w.en + ND, 1, D, w.S_ga, N, 0, N, N, D, nullptr, nullptr, 1};
        if ((rc = run_gemm(g, use_tc, sc_ga, s_ga))) return rc;
    }
    align_rows_kernel<<<N, 128, 0, s_ga>>>(w.S_ga, labels, hp->alpha, hp->beta, hp->scale_pos, hp->scale_neg, N, w.rows_ga, grads);
    TRB_LAUNCH_OK();
    if (grads) {
        GemmArgs gv{w.S_ga, N, 1, w.en + ND, D, 1, w.dq_ga, D, 0, N, D, N, nullptr, nullptr, 1};
        if ((rc = run_gemm(gv, use_tc, sc_ga, s_ga))) return rc;
        GemmArgs gt{w.S_ga, 1, N, w.en, D, 1, w.dq_ga + ND, D, 0, N, D, N, nullptr, nullptr, 1};
        if ((rc = run_gemm(gt, use_tc, sc_ga, s_ga))) return rc;
        normalize_backward_kernel<<<(rows + 7) / 8, 256, 0, s_ga>>>(w.dq_ga, nullptr, nullptr, nullptr, w.en, w.inv_e, d_ga, N, D);
        TRB_LAUNCH_OK();
    }
    }
    TRB_CUDA_OK(cudaEventRecord(fk->j2, s_ga));

    // ---- instance branch (caller's stream): column norms, logits of both modalities at once (losses.py:51-54),
    //      label-smoothed CE rows, dE = dZ @ What^T (split over classes), dWhat = E^T @ dZ, column-normalisation Jacobian
    if (!(roles & 1)) {
    column_inv_norm_kernel<<<(C + 31) / 32, 256, 0, st>>>(projection, w.inv_c, D, C);
    TRB_LAUNCH_OK();
    {
        GemmArgs g{w.E2, D, 1, projection, C, 1, w.Z, C, 0, rows, C, D, nullptr, w.inv_c, 1};
        if ((rc = run_gemm(g, use_tc, sc_inst, st))) return rc;
    }
    instance_rows_kernel<<<rows, 256, 0, st>>>(w.Z, labels, hp->epsilon, N, C, w.rows_inst, grads);
    TRB_LAUNCH_OK();
    if (grads) {
        // the two contractions with dZ are independent: dWhat goes to helper stream 3
        if (d_projection) {
            TRB_CUDA_OK(cudaEventRecord(fk->fork3, st));
            TRB_CUDA_OK(cudaStreamWaitEvent(fk->s3, fk->fork3, 0));
            GemmArgs g2{w.E2, 1, D, w.Z, C, 1, d_projection, C, 0, D, C, rows, nullptr, nullptr, 1};
            if ((rc = run_gemm(g2, use_tc, sc_dw, fk->s3))) return rc;
            projection_backward_kernel<<<(C + 31) / 32, 256, 0, fk->s3>>>(d_projection, projection, w.inv_c, D, C);
            TRB_LAUNCH_OK();
            TRB_CUDA_OK(cudaEventRecord(fk->j3, fk->s3));
        }
        GemmArgs g{w.Z, C, 1, projection, 1, C, w.part, D, 2 * ND, rows, D, C, w.inv_c, nullptr, split_inst};
        if ((rc = run_gemm(g, use_tc, sc_inst, st))) return rc;
        reduce_partials_kernel<<<(unsigned)((2 * ND + 255) / 256), 256, 0, st>>>(d_inst, w.part, split_inst, 2 * ND);
        TRB_LAUNCH_OK();
        if (d_projection) TRB_CUDA_OK(cudaStreamWaitEvent(st, fk->j3, 0));
    }
    }

    // ---- join, then the three loss scalars in a fixed reduction order
    TRB_CUDA_OK(cudaStreamWaitEvent(st, fk->j1, 0));
    TRB_CUDA_OK(cudaStreamWaitEvent(st, fk->j2, 0));
    loss_reduce_kernel<<<1, 256, 0, st>>>(w.rows_inst, w.rows_nce, w.rows_ga, N, losses);
    TRB_LAUNCH_OK();
    if (queue_ptr != nullptr)      // every branch has been joined back: the queues are no longer read
        return trb_enqueue(const_cast<float*>(v_queue), const_cast<float*>(t_queue), const_cast<int64_t*>(id_queue), queue_ptr,
                           v_key_n, t_key_n, labels, N, D, K, (trb_stream_t)st);
    return 0;
}

#define TRB_LOSS_ARGS                                                                                                          \
    const float *v_embed, const float *t_embed, const float *v_qraw, const float *t_qraw, const float *v_key,                  \
        const float *t_key, int normalize_keys, float *v_key_n, float *t_key_n, const int64_t *labels, const float *v_queue,   \
        const float *t_queue, const int64_t *id_queue, const float *projection, const trb_moco_shape *shape,                    \
        const trb_moco_hparams *hp, float *losses, float *d_inst, float *d_nce, float *d_ga, float *d_projection,              \
        void *workspace, int64_t workspace_bytes, cudaStream_t st, int64_t *queue_ptr
#define TRB_LOSS_PASS                                                                                                          \
    v_embed, t_embed, v_qraw, t_qraw, v_key, t_key, normalize_keys, v_key_n, t_key_n, labels, v_queue, t_queue, id_queue,      \
        projection, shape, hp, losses, d_inst, d_nce, d_ga, d_projection, workspace, workspace_bytes, st

int trb_moco_loss_f32(TRB_LOSS_ARGS) { return moco_loss_impl(TRB_LOSS_PASS, false, queue_ptr); }
int trb_moco_loss_tc(TRB_LOSS_ARGS) { return moco_loss_impl(TRB_LOSS_PASS, true, queue_ptr); }

// launches of the enqueue that follows the loss sequence: 0 when it rides inside the fused cooperative kernel
int trb_moco_step_extra_launches_impl(const trb_moco_shape* s, int precision) {
    if (precision == 1 && fused_loss_supported(s->N, s->D, s->K, s->C, sm_count())) {
        const char* e = getenv("TRB_FUSED_ROLES");
        if (e == nullptr || (atoi(e) & 7) == 7) return 0;
    }
    return 2;
}

// debug read-back from the caller's workspace (fused bf16 path only)
int trb_moco_loss_debug_impl(const void* workspace, const trb_moco_shape* s, int what, void* host_out) {
    if (!fused_loss_supported(s->N, s->D, s->K, s->C, sm_count()) && !fused_windows_supported(s->N, s->D, s->K, s->C, sm_count())) {
        trb_set_error("moco_loss debug: the shape does not take the fused path");
        return TRB_ERR_UNSUPPORTED;
    }
    const Workspace w = carve(const_cast<void*>(workspace), s->N, s->D, s->K, s->C, true);
    return fused_loss_debug_copy(w.fused, s->N, s->D, s->K, s->C, what, host_out);
}
