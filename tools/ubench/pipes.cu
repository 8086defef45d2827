// Micro-benchmark of issue / pipe throughput for the instruction mixes the stream epilogue can count with (sm_100a).
// Each kernel runs ITER iterations of a fully unrolled body on 32 independent values per thread; 4 warps per SMSP.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <cuda_fp16.h>
#define ITER 2000
__device__ __forceinline__ void add2(unsigned long long& acc, float a, float b) {
    unsigned long long t;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(a), "f"(b));
    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(t));
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ in, float* out, float c, long long* cycles) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 1023];
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    float f0 = 0, f1 = 0, f2 = 0, f3 = 0;
    unsigned long long a0 = 0, a1 = 0;
    const float H = 1.2676506002282294e30f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        const float te = c + (float)it;          // changes per iteration: nothing hoists
        const float cc = -te * H;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (MODE == 0) {        // FADD + LEA.HI
                c0 += __float_as_uint(te - v[j]) >> 31; c1 += __float_as_uint(te - v[j + 1]) >> 31;
                c2 += __float_as_uint(te - v[j + 2]) >> 31; c3 += __float_as_uint(te - v[j + 3]) >> 31;
            } else if (MODE == 1) { // FFMA.SAT + FADD
                f0 += __saturatef(fmaf(v[j], H, cc)); f1 += __saturatef(fmaf(v[j + 1], H, cc));
                f2 += __saturatef(fmaf(v[j + 2], H, cc)); f3 += __saturatef(fmaf(v[j + 3], H, cc));
            } else if (MODE == 2) { // 2 FFMA.SAT + FADD2
                add2(a0, __saturatef(fmaf(v[j], H, cc)), __saturatef(fmaf(v[j + 1], H, cc)));
                add2(a1, __saturatef(fmaf(v[j + 2], H, cc)), __saturatef(fmaf(v[j + 3], H, cc)));
            } else if (MODE == 3) { // half/half mix of 0 and 1
                c0 += __float_as_uint(te - v[j]) >> 31; c1 += __float_as_uint(te - v[j + 1]) >> 31;
                f2 += __saturatef(fmaf(v[j + 2], H, cc)); f3 += __saturatef(fmaf(v[j + 3], H, cc));
            } else if (MODE == 4) { // mix: 2 pairs FADD2-type, 2 pairs LEA-type
                add2(a0, __saturatef(fmaf(v[j], H, cc)), __saturatef(fmaf(v[j + 1], H, cc)));
                c2 += __float_as_uint(te - v[j + 2]) >> 31; c3 += __float_as_uint(te - v[j + 3]) >> 31;
            } else if (MODE == 5) { // FADD only (FMA-pipe rate of a 2-register add)
                f0 += v[j] + te; f1 += v[j + 1]; f2 += v[j + 2]; f3 += v[j + 3];
            } else if (MODE == 6) { // FFMA.SAT only
                f0 = __saturatef(fmaf(v[j], H, f0)); f1 = __saturatef(fmaf(v[j + 1], H, f1));
                f2 = __saturatef(fmaf(v[j + 2], H, f2)); f3 = __saturatef(fmaf(v[j + 3], H, f3));
            } else if (MODE == 7) { // FADD2 only
                add2(a0, v[j], v[j + 1]); add2(a1, v[j + 2], v[j + 3]);
            } else if (MODE == 9) { // FFMA.SAT x4 + integer adds of the 1.0f bit patterns: 3 per IADD3, one shifted accumulate (>> 23 = 127 per hit) per group
                const unsigned i0 = __float_as_uint(__saturatef(fmaf(v[j], H, cc))), i1 = __float_as_uint(__saturatef(fmaf(v[j + 1], H, cc)));
                const unsigned i2 = __float_as_uint(__saturatef(fmaf(v[j + 2], H, cc))), i3 = __float_as_uint(__saturatef(fmaf(v[j + 3], H, cc)));
                c0 += (i0 + i1 + i2 + i3) >> 23;          // 4 x 0x3F800000 = 0xFE000000 still fits
            } else if (MODE == 8) { // FFMA.SAT + IADD-style accumulate of the indicator bits (ALU)  (acc += bits >> 29)
                c0 += __float_as_uint(__saturatef(fmaf(v[j], H, cc))) >> 29; c1 += __float_as_uint(__saturatef(fmaf(v[j + 1], H, cc))) >> 29;
                c2 += __float_as_uint(__saturatef(fmaf(v[j + 2], H, cc))) >> 29; c3 += __float_as_uint(__saturatef(fmaf(v[j + 3], H, cc))) >> 29;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(c0 + c1 + c2 + c3) + f0 + f1 + f2 + f3 + __uint_as_float((unsigned)a0) + __uint_as_float((unsigned)(a0 >> 32)) + __uint_as_float((unsigned)a1) + __uint_as_float((unsigned)(a1 >> 32));
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
// Groups of G indicators summed as integers (IADD3 takes three), then ONE shifted accumulate per group.
template <int G>
__global__ void __launch_bounds__(512, 1) kg(const float* __restrict__ in, float* out, float c, long long* cycles) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 1023];
    unsigned c0 = 0, c1 = 0;
    const float H = 1.2676506002282294e30f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        const float te = c + (float)it;
        const float cc = -te * H;
        unsigned ind[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) ind[j] = __float_as_uint(__saturatef(fmaf(v[j], H, cc)));
        int g = 0;
#pragma unroll
        for (int j = 0; j < 32; j += G) {
            unsigned a = 0;
#pragma unroll
            for (int i = 0; i < G; ++i) if (j + i < 32) a += ind[j + i];
            if (g & 1) c1 += a >> 23; else c0 += a >> 23;
            ++g;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(c0 + c1);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int G> void rung(const float* in, float* out, long long* cyc, int threads) {
    for (int rep = 0; rep < 2; ++rep) { kg<G><<<148, threads>>>(in, out, 0.5f, cyc); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32.0 / 4.0;
    printf("FFMA.SAT + integer group sums of %d + one shifted accumulate, %2.0f warps per SMSP: %8lld cycles  %.3f SMSP-cycles per (value,threshold) warp-op\n",
           G, warps_per_smsp, h, (double)h / (ITER * 32.0 * warps_per_smsp));
}
// Four thresholds per value like the stream epilogue: NF of them counted on the FMA pipe (2 x FFMA.SAT + FADD2 per value pair),
// the rest with a packed difference (FADD2 te2 - v2) and two LEA.HI sign-bit accumulates on the ALU pipe.
__device__ __forceinline__ void sub2(float& d0, float& d1, float te, float v0, float v1) {
    unsigned long long a, b;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(te));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(-v0), "f"(-v1));
    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(a));
}
template <int NF>
__global__ void __launch_bounds__(512, 1) k4(const float* __restrict__ in, float* out, float c, long long* cycles) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 1023];
    unsigned cnt[4] = {0, 0, 0, 0};
    unsigned long long acc[4] = {0, 0, 0, 0};
    const float H = 1.2676506002282294e30f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float te = c + (float)(it + r);
            const float cc = -te * H;
            if (r < NF) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) add2(acc[r], __saturatef(fmaf(v[j], H, cc)), __saturatef(fmaf(v[j + 1], H, cc)));
            } else {
                unsigned c0 = 0, c1 = 0;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float d0, d1;
                    sub2(d0, d1, te, v[j], v[j + 1]);
                    c0 += __float_as_uint(d0) >> 31; c1 += __float_as_uint(d1) >> 31;
                }
                cnt[r] += c0 + c1;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int r = 0; r < 4; ++r) s += (float)cnt[r] + __uint_as_float((unsigned)acc[r]) + __uint_as_float((unsigned)(acc[r] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
// Four thresholds, the shipped formulation: FFMA.SAT indicators, integer sums of three (IADD3), one shifted accumulate (LEA.HI)
template <int G>
__global__ void __launch_bounds__(512, 1) k4i(const float* __restrict__ in, float* out, float c, long long* cycles) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 1023];
    unsigned cnt[4] = {0, 0, 0, 0};
    const float H = 1.2676506002282294e30f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float cc = -(c + (float)(it + r)) * H;
            unsigned acc = cnt[r];
#pragma unroll
            for (int j = 0; j < 32; j += G) {
                unsigned a = 0;
#pragma unroll
                for (int i = 0; i < G; ++i) if (j + i < 32) a += __float_as_uint(__saturatef(fmaf(v[j + i], H, cc)));
                acc += a >> 23;
            }
            cnt[r] = acc;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(cnt[0] + cnt[1] + cnt[2] + cnt[3]);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
// the same with the indicators as volatile asm: ptxas keeps them in source order (threshold-major), so every FFMA.SAT of a
// threshold finds its compare value in the operand reuse cache
__device__ __forceinline__ unsigned ind_v(float v, float cc) {
    float r;
    asm volatile("fma.rn.sat.f32 %0, %1, 0f71800000, %2;" : "=f"(r) : "f"(v), "f"(cc));
    return __float_as_uint(r);
}
template <int G>
__global__ void __launch_bounds__(512, 1) k4v(const float* __restrict__ in, float* out, float c, long long* cycles) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 1023];
    unsigned cnt[4] = {0, 0, 0, 0};
    const float H = 1.2676506002282294e30f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float cc = -(c + (float)(it + r)) * H;
            unsigned acc = cnt[r];
#pragma unroll
            for (int j = 0; j < 32; j += G) {
                unsigned a = 0;
#pragma unroll
                for (int i = 0; i < G; ++i) if (j + i < 32) a += ind_v(v[j + i], cc);
                acc += a >> 23;
            }
            cnt[r] = acc;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(cnt[0] + cnt[1] + cnt[2] + cnt[3]);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int G> void run4v(const float* in, float* out, long long* cyc, int threads) {
    for (int rep = 0; rep < 2; ++rep) { k4v<G><<<148, threads>>>(in, out, 0.5f, cyc); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32.0 / 4.0;
    printf("4 thresholds, volatile FFMA.SAT + IADD3 sums of %d + LEA.HI, %2.0f warps per SMSP: %8lld cycles  %.3f SMSP-cycles per (value,threshold) warp-op\n",
           G, warps_per_smsp, h, (double)h / (ITER * 32.0 * 4.0 * warps_per_smsp));
}
template <int G> void run4i(const float* in, float* out, long long* cyc, int threads) {
    for (int rep = 0; rep < 2; ++rep) { k4i<G><<<148, threads>>>(in, out, 0.5f, cyc); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32.0 / 4.0;
    printf("4 thresholds, FFMA.SAT + IADD3 sums of %d + LEA.HI, %2.0f warps per SMSP: %8lld cycles  %.3f SMSP-cycles per (value,threshold) warp-op\n",
           G, warps_per_smsp, h, (double)h / (ITER * 32.0 * 4.0 * warps_per_smsp));
}
// Four thresholds counted in packed half precision: the 32 values are rounded toward zero to 16 half2 pairs once (F2FP), then per
// threshold one HFMA2.SAT (indicator: 0 below, 0.5 equal, 1 above) and one HADD2 per PAIR of values.  FLUSH = also fold the four
// half2 sums into integer counters and test them for a fractional part once per 32 values, like the epilogue would.
template <int FLUSH>
__global__ void __launch_bounds__(512, 1) k4h(const float* __restrict__ in, float* out, float c, long long* cycles) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 1023];
    unsigned cnt[4] = {0, 0, 0, 0};
    unsigned amb = 0;
    __half2 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = __float2half2_rn(0.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        __half2 hv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            unsigned u;
            asm("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[2 * j + 1] + (float)it), "f"(v[2 * j]));
            hv[j] = *reinterpret_cast<__half2*>(&u);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const __half2 sc = __float2half2_rn(512.f + (float)r);
            const __half2 cc = __float2half2_rn(0.5f - (c + (float)(it + r)) * 512.f);
            __half2 a = FLUSH ? __float2half2_rn(0.f) : acc[r];
#pragma unroll
            for (int j = 0; j < 16; ++j) a = __hadd2(a, __hfma2_sat(hv[j], sc, cc));
            if (FLUSH) {
                const float t = __low2float(a) + __high2float(a);
                const int ti = (int)t;
                amb |= (t != (float)ti);
                cnt[r] += ti;
            } else acc[r] = a;
        }
    }
    long long t1 = clock64();
    float sres = (float)(cnt[0] + cnt[1] + cnt[2] + cnt[3] + amb);
    for (int r = 0; r < 4; ++r) sres += __low2float(acc[r]) + __high2float(acc[r]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = sres;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int FLUSH> void run4h(const float* in, float* out, long long* cyc, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k4h<FLUSH><<<148, threads>>>(in, out, 0.5f, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0); k4h<FLUSH><<<148, threads>>>(in, out, 0.5f, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32.0 / 4.0;
    printf("4 thresholds in half2 (F2FP once, HFMA2.SAT + HADD2 per pair)%s, %2.0f warps per SMSP: %8lld cycles  %.3f SMSP-cycles per (value,threshold) warp-op  [%.3f ms whole kernel]\n",
           FLUSH ? " + per-chunk flush" : "", warps_per_smsp, h, (double)h / (ITER * 32.0 * 4.0 * warps_per_smsp), ms);
}
template <int G> void run4e(const float* in, float* out, long long* cyc, int threads) {   // event-timed reference: the IADD3 formulation
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k4i<G><<<148, threads>>>(in, out, 0.5f, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0); k4i<G><<<148, threads>>>(in, out, 0.5f, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("   (reference, event-timed: FFMA.SAT + IADD3 sums of %d, %d threads: %.3f ms whole kernel)\n", G, threads, ms);
}
template <int NF> void run4(const float* in, float* out, long long* cyc, int threads) {
    for (int rep = 0; rep < 2; ++rep) { k4<NF><<<148, threads>>>(in, out, 0.5f, cyc); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32.0 / 4.0;
    printf("4 thresholds, %d on the FMA pipe / %d on the ALU pipe, %2.0f warps per SMSP: %8lld cycles  %.3f SMSP-cycles per (value,threshold) warp-op\n",
           NF, 4 - NF, warps_per_smsp, h, (double)h / (ITER * 32.0 * 4.0 * warps_per_smsp));
}
template <int MODE> void run(const char* name, const float* in, float* out, long long* cyc) {
    k<MODE><<<148, 512>>>(in, out, 0.5f, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, 512>>>(in, out, 0.5f, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // 16 warps / 4 SMSP = 4 warps per scheduler; pairs per warp = ITER*32
    double per_pair = (double)h / (ITER * 32.0 * 4.0);
    printf("%-44s %8lld cycles  %.3f SMSP-cycles per (value,threshold) warp-op\n", name, h, per_pair);
}
int main() {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (float)i / 1024.f;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    run<0>("FADD + LEA.HI", in, out, cyc);
    run<1>("FFMA.SAT + FADD", in, out, cyc);
    run<2>("2 FFMA.SAT + FADD2", in, out, cyc);
    run<3>("mix 1:1 (FADD+LEA.HI | FFMA.SAT+FADD)", in, out, cyc);
    run<4>("mix 1:1 (FFMA.SAT x2+FADD2 | FADD+LEA.HI x2)", in, out, cyc);
    run<5>("FADD only (2 per slot... see code)", in, out, cyc);
    run<6>("FFMA.SAT only", in, out, cyc);
    run<7>("FADD2 only (1 per 2 values)", in, out, cyc);
    run<8>("FFMA.SAT + LEA.HI(bits>>29)", in, out, cyc);
    run<9>("4 FFMA.SAT + 2 IADD3 + LEA.HI(>>23)", in, out, cyc);
    for (int threads : {256, 512}) { rung<2>(in, out, cyc, threads); rung<3>(in, out, cyc, threads); rung<4>(in, out, cyc, threads); }
    for (int threads : {256, 512}) { run4h<0>(in, out, cyc, threads); run4h<1>(in, out, cyc, threads); run4e<3>(in, out, cyc, threads); }
    for (int threads : {128, 256, 512}) { run4v<3>(in, out, cyc, threads); run4v<4>(in, out, cyc, threads); }
    for (int threads : {128, 256, 512}) { run4i<3>(in, out, cyc, threads); run4i<4>(in, out, cyc, threads); }
    for (int threads : {128, 256, 512}) {
        run4<4>(in, out, cyc, threads); run4<3>(in, out, cyc, threads); run4<2>(in, out, cyc, threads);
        run4<1>(in, out, cyc, threads); run4<0>(in, out, cyc, threads);
    }
    return 0;
}
