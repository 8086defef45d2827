// bf16 tensor-core retrieval stream for sm_100a: tcgen05.mma (M=128, N=256, K=16, fp32 accumulate in
// TMEM), operands staged into shared memory by cp.async.bulk (TMA engine) from the packed
// pre-swizzled HBM layout of tc_common.cuh, a 5-stage mbarrier ring (D = 256), double-buffered TMEM
// accumulators, and an epilogue that consumes the similarity tile straight out of TMEM:
// per-query top-10 and exact rank counts.  The [Q,G] similarity matrix never exists in HBM.
// What bounds it and every build-time knob below (TRB_TC_*): profiles/r02_stream_timeline.md,
// profiles/r02_stream_experiments.md.
//
// Replaces (with evaluation.py:117-120 fused in by trb_pack_rows_bf16) the reference's
//   similarity = text @ image.T ; argsort ; matches ; cumsum        lib/data/metrics/evaluation.py:11-37,120
//
// Warp roles (2 control + 8 epilogue warps = 320 threads, one persistent CTA per SM):
//   warp 0        : TMEM allocator / deallocator; lane 0 = producer - bulk copies: query tile (resident per work unit)
//                   + gallery k-chunk ring
//   warp 1 lane 0 : MMA issuer - tcgen05.mma into TMEM buffer t&1, tcgen05.commit -> barriers
//   warps 2..9    : epilogue  - warp w reads TMEM lanes 32*(w%4).. (its 32 query rows) and the
//                   128-column slice (w-2)/4 of the 256-column accumulator, 32 columns at a time
#include "tc_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

using namespace tc;

constexpr int TILE_M = 128;          // queries per CTA tile (TMEM lanes)
// Gallery rows per MMA tile = TMEM columns per accumulator buffer.  The 512 TMEM columns hold 512 / TILE_N accumulators: a ring of
// FOUR 128-column tiles instead of two 256-column ones lets the MMAs run up to three tiles ahead of the slowest epilogue warp (with
// two buffers the MMAs of tile t + 2 wait for the last warp's release of tile t: per-tile time stamps, profiles/r02_stream_timeline.md)
#ifndef TRB_TC_TILE_N
#define TRB_TC_TILE_N 256
#endif
constexpr int TILE_N = TRB_TC_TILE_N;
static_assert(TILE_N == 128 || TILE_N == 256, "TILE_N: one or two 128-row gallery blocks");
constexpr int NTB = 512 / TILE_N;    // accumulator buffers in TMEM
constexpr int PACK_ROWS = 256;       // row padding of the packed operands (independent of the tile width)
constexpr int UMMA_K = 16;
constexpr int TILE_BLOCKS = TILE_N / 128;
constexpr int STAGE_BYTES = TILE_BLOCKS * BLOCK_BYTES;   // TILE_N gallery rows x 64 k  = 16 / 32 KiB
#ifndef TRB_EPI_WARPS
#define TRB_EPI_WARPS 8
#endif
// two control warps (producer + TMEM allocator, MMA issuer) and the epilogue warps.  8 epilogue warps keep the whole row
// state in registers (167, no spills); 16 warps at the 96-register cap spill inside the hot loop and measured slower
// (71.7 vs 57.1 ms per 100k x 1M stream, gpurun_out/tc_ab.log of round 2)
constexpr int EPI_WARP0 = 2;
constexpr int NUM_THREADS = 32 * EPI_WARP0 + 32 * TRB_EPI_WARPS;
constexpr int NUM_EPI_WARPS = TRB_EPI_WARPS;
constexpr int NUM_COLGRP = NUM_EPI_WARPS / 4;                 // epilogue warps per TMEM lane quarter (= per SM sub-partition)
#ifndef TRB_TC_CHUNK
#define TRB_TC_CHUNK 32
#endif
constexpr int CH = TRB_TC_CHUNK;                               // accumulator columns a thread holds in registers at a time (16 | 32)
constexpr int TILE_CHUNKS = TILE_N / CH;
constexpr bool COLGRP_CONTIGUOUS = TILE_CHUNKS % NUM_COLGRP == 0;
// column group cg takes a contiguous run of chunks when the groups divide the tile evenly, otherwise chunks cg, cg + groups, ...
// (3 groups over 8 chunks: 3 + 3 + 2; TRB_EPI_WARPS=12 measured 55.8 vs 56.0 ms: the epilogue is throughput bound, not latency
// bound); either way a warp walks its chunks in ascending gallery order
__host__ __device__ constexpr int chunks_of(int cg) {
    return COLGRP_CONTIGUOUS ? TILE_CHUNKS / NUM_COLGRP : (TILE_CHUNKS - cg + NUM_COLGRP - 1) / NUM_COLGRP;
}
__host__ __device__ constexpr int chunk_col(int cg, int i) {
    return (COLGRP_CONTIGUOUS ? cg * (TILE_CHUNKS / NUM_COLGRP) + i : i * NUM_COLGRP + cg) * CH;
}
constexpr int LISTS_PER_SPLIT = NUM_COLGRP;                   // candidate lists a query gets per gallery split
constexpr int MAX_STAGES = 16;
// Build-time tuning knobs (A/B builds through textreid_b200.build.build_variant):
//   TRB_TC_SKIP_R   : thresholds per row (sorted descending) that get a warp-uniform "no value of this chunk reaches it" test
//   TRB_TC_HOT_SPIN : 1 = the producer / MMA threads poll their mbarriers in a hot loop (round-1 behaviour)
//   TRB_TC_COUNT_FMA: 1 = FFMA.SAT indicator + packed fp32 add (FADD2), all on the FMA pipe (shipped);
//                     2 = FFMA.SAT indicator + INTEGER sums of the 1.0f bit patterns on the ALU pipe: three indicators per IADD3,
//                         one shifted accumulate (LEA.HI, >> 23 = 127 per hit) per group -- 5 instructions per 3 (value, threshold)
//                         pairs on two pipes; measured 59.6 vs 57.5 ms (the micro-benchmark puts every formulation at 2.2 - 2.4
//                         SM cycles per pair: profiles/r02_stream_experiments.md);
//                     0 = FADD + LEA.HI (FMA + ALU pipe)
#ifndef TRB_TC_SKIP_R
#define TRB_TC_SKIP_R 0
#endif
#ifndef TRB_TC_COUNT_FMA
#define TRB_TC_COUNT_FMA 1
#endif
//   TRB_TC_PREFETCH : 1 = the epilogue keeps the TMEM load of the next chunk in flight while it consumes the current one
//                     (measured slower on B200, 59.4 vs 56.3 ms: the second register buffer costs more than the latency it hides)
#ifndef TRB_TC_PREFETCH
#define TRB_TC_PREFETCH 0
#endif

#ifndef TRB_TC_HOT_SPIN
#define TRB_TC_HOT_SPIN 0
#endif
//   TRB_TC_COUNT_FIRST: 1 = the register counts of a chunk are taken BEFORE the rare-case vote is consumed (its four dependent
//                       instructions overlap the arithmetic); the rare path then corrects instead of preparing: a threshold that
//                       switches to its strict compare in this chunk takes back the ties at / after its item, the padded tile
//                       takes back its zero rows
#ifndef TRB_TC_COUNT_FIRST
#define TRB_TC_COUNT_FIRST 1
#endif
//   TRB_TC_FAST_CHUNK: 1 = one warp-uniform test per chunk ("nothing rare in any lane") in front of the per-condition branches
#ifndef TRB_TC_FAST_CHUNK
#define TRB_TC_FAST_CHUNK 1
#endif
constexpr int SMEM_MAX = 232448;      // 227 KiB opt-in limit per CTA on sm_100

struct Params {
    const uint8_t* q_packed;
    const uint8_t* g_packed;
    int64_t Q, G;            // valid (unpadded) row counts
    int kchunks;             // D / 64
    int nstages;
    const int64_t* q_row_id; // [Qp] original query number per packed row, -1 = padding
    const int64_t* g_row_id; // mode 1: [Gp] global gallery index per (pid-sorted) packed row, -1 = padding
    int64_t g_base;          // mode 0: the gallery is packed in index order; global index = g_base + packed row
    const int64_t* rel_ptr;  // [Qorig+1]
    float* thr;              // [total]  (mode 0: read, mode 1: written)
    int64_t* thr_gidx;       // [total]
    const int32_t* band_lo;  // mode 1, per packed query row
    const int32_t* band_hi;
    const int32_t* rel_off;
    int nsplit;
    float* cand_sim;         // [Qorig, LISTS_PER_SPLIT*nsplit, 10]
    int64_t* cand_idx;
    int32_t* cnt;            // [total]
    int64_t num_qtiles, num_gtiles, num_units;
    int64_t full_qtiles;     // mode 0: query tiles [0, full_qtiles) stream the whole gallery in one unit (whole waves of the
                             // persistent grid); the remaining tiles are cut into nsplit gallery pieces to balance the last wave
    uint32_t wait_hint_ns;   // suspend-time hint of the epilogue warps' accumulator waits
    int debug;               // builds with -DTRB_TC_PROBE only (TRB_TC_DEBUG env): 1 = epilogue skips the arithmetic, 2 = and the TMEM read,
                             // 4 = the producer stops loading gallery tiles once the ring is full, 8 = no top-10 maintenance
};

// Waits of the two single-thread roles.  Polling costs issue slots on the scheduler that also runs a quarter of the epilogue
// warps: in the round-2 profile the try_wait loops of these two threads were 12 % of all issued instructions (try_wait with a
// suspend hint compiles to TRYWAIT + NANOSLEEP.SYNCS, which wakes on every mbarrier event of the CTA, i.e. every ~16 ns here).
// A plain nanosleep really parks the thread; the double-buffered accumulator and the multi-stage ring give both roles a full
// tile / several stages of slack, so a wake-up granularity of `ns` costs nothing.
__device__ __forceinline__ void role_wait(uint64_t* bar, uint32_t parity, uint32_t ns) {
#if TRB_TC_HOT_SPIN
    mbar_wait(bar, parity);
#else
#pragma unroll 1
    for (int i = 0; i < 4; ++i)
        if (mbar_try_wait(bar, parity)) return;
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (++n > (1u << 24)) {
            printf("trb: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
#endif
}

#ifdef TRB_TC_PROBE
// probe: per-tile time stamps of CTA 0 (TRB_TC_DEBUG & 16): [0] MMA issuer got the accumulator back (t_empty), [1] MMA issuer
// committed the tile, [2] epilogue warp 2 saw the tile complete (t_full), [3] epilogue warp 2 released it,
// [4] epilogue warp 6 saw it, [5] epilogue warp 6 released it
constexpr int STAMP_TILES = 48, STAMP_FIRST = 1000;
__device__ unsigned long long g_stamps[16][STAMP_TILES];   // [6 + w]: epilogue warp 2 + w released the tile (w < 8); [14] MMA issuer starts to wait for the accumulator
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRB_STAMP(row, t)                                                                                              \
    do {                                                                                                               \
        if ((p.debug & 16) && blockIdx.x == 0 && (t) >= STAMP_FIRST && (t) < STAMP_FIRST + STAMP_TILES)                 \
            g_stamps[row][(t) - STAMP_FIRST] = gtime();                                                                \
    } while (0)
#else
#define TRB_STAMP(row, t) do { } while (0)
#endif

// tail flag of a tile: a compile-time constant (std::true_type / std::false_type) or a run-time value
template <class T>
__device__ __forceinline__ constexpr uint32_t tail_value(T v) {
    if constexpr (std::is_same<T, uint32_t>::value) return v;
    else return T::value ? 1u : 0u;
}

struct UnitInfo {
    int64_t qt, split, t_lo, t_hi;
};

template <int MODE>
__device__ __forceinline__ UnitInfo unit_info(const Params& p, int64_t u) {
    UnitInfo ui;
    if (MODE == 0) {
        if (u < p.full_qtiles) {
            ui.split = -1;                   // unsplit: writes list group 0 and pads the others
            ui.qt = u;
            ui.t_lo = 0;
            ui.t_hi = p.num_gtiles;
        } else {
            const int64_t r = u - p.full_qtiles, rem = p.num_qtiles - p.full_qtiles;
            ui.split = r / rem;              // split-major: concurrent CTAs stream the same gallery range
            ui.qt = p.full_qtiles + r % rem;
            ui.t_lo = p.num_gtiles * ui.split / p.nsplit;
            ui.t_hi = p.num_gtiles * (ui.split + 1) / p.nsplit;
        }
    } else {
        ui.split = 0;
        ui.qt = u;
        const int64_t r0 = ui.qt * TILE_M;
        const int64_t r1 = min(r0 + TILE_M, p.Q) - 1;
        const int64_t lo = p.band_lo[r0], hi = p.band_hi[r1];   // both monotone in the pid-sorted row order
        ui.t_lo = lo / TILE_N;
        ui.t_hi = hi > lo ? (hi + TILE_N - 1) / TILE_N : ui.t_lo;
    }
    return ui;
}

// ---------------------------------------------------------------------------------------------
// epilogue: per query row (= one TMEM lane = one thread) state and the per-chunk consumer
//
// The stream (mode 0) walks the gallery in GLOBAL INDEX order.  The pinned ranking order is (similarity desc, index asc),
// so for a relevant item r with similarity thr: an item g ranks before r  <=>  s_g > thr, or s_g == thr and g < r
//   <=>  s_g >= thr for g < r  and  s_g > thr for g > r.
// Chunks that lie entirely before r therefore count against nextbelow(thr) (">=" as a strict compare), chunks after r
// against thr itself, and only the one chunk that contains r needs a correction.  No per-value tie detection, no index
// look-ups: 2 instructions per (value, threshold): d = t - v (FADD), count += bits(d) >> 31 (LEA.HI).
// ---------------------------------------------------------------------------------------------
// acc.lo += a, acc.hi += b in one instruction (packed fp32 add, SASS FADD2)
__device__ __forceinline__ void add2(unsigned long long& acc, float a, float b) {
    unsigned long long t;
    asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(a), "f"(b));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(t));
}
constexpr float COUNT_SCALE = 1.2676506002282294e30f;     // 2^100
constexpr float COUNT_MIN_ABS = 1.734723475976807e-18f;    // 2^-59
__device__ __forceinline__ float next_below(float x) {        // largest float < x (x finite or +inf)
    const uint32_t b = __float_as_uint(x);
    const uint32_t r = (x > 0.0f) ? b - 1u : (((b << 1) == 0u) ? 0x80000001u : b + 1u);
    return __uint_as_float(r);
}

// v[j] for a run-time j without spilling v to local memory: select tree (CH - 1 FSEL)
__device__ __forceinline__ float select_lane(const float (&v)[CH], int j) {
    float a[CH / 2], b[CH / 4], c[CH / 8];
    const bool b0 = j & 1, b1 = j & 2, b2 = j & 4, b3 = j & 8, b4 = j & 16;
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) a[i] = b0 ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) b[i] = b1 ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < CH / 8; ++i) c[i] = b2 ? b[2 * i + 1] : b[2 * i];
    const float d0 = b3 ? c[1] : c[0];
    if (CH == 16) return d0;
    const float d1 = b3 ? c[CH / 8 - 1] : c[CH / 8 - 2];
    return b4 ? d1 : d0;
}

// TMEM -> registers, split into issue and completion so that a load can be in flight under arithmetic.  The wait takes the
// destination registers as read-write operands: every later use of them depends on it.
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(float (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]),
                   "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]),
                   "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}

template <int RTN>
struct RowState {
    float ts[TRB_TOPK];      // best-first similarities
    int tr[TRB_TOPK];        // their packed (= local) gallery rows
    // Register-resident thresholds, sorted by similarity DESCENDING (unused slots = +inf sort first), so that slot r of
    // every lane of a warp is "the r-th best relevant item of my row" and the warp-uniform skip test below fires together.
    // te: the value the stream compares against -- nextbelow(thr) until the stream reaches the item (ties before it count),
    // thr itself afterwards.  With TRB_TC_COUNT_FMA the register holds -te * 2^100 instead: sat(v * 2^100 - te * 2^100) is
    // exactly 1 for v > te and 0 otherwise as long as |te| >= 2^-59 (the smallest positive difference, one ulp >= 2^-83, still
    // scales to >= 1); thresholds closer to zero than that take the exact slow path (`slow`, CSR positions).
    float te[RTN];
    int sw[RTN];             // first chunk start g0 at which te switches to thr (= local row of the item - (CH - 1)); INT32_MAX = done
#if TRB_TC_COUNT_FMA == 2
    uint32_t ci[RTN];             // 127 x hits: (k x 0x3F800000) >> 23 = 127 k for the k <= 3 indicators of a group
#elif TRB_TC_COUNT_FMA
    unsigned long long cf[RTN];   // two fp32 counters per slot (even / odd values of a chunk), bumped by one FADD2 per value pair
#else
    int cnt[RTN];
#endif
    int next_sw;             // min over sw[]
    uint32_t perm;           // 4 bits per register slot: position of the item inside the row's CSR range
    uint32_t slow;           // bit i: CSR position i (< RTN) is counted by the exact slow path, not in registers
    int64_t s_lo, s_hi;      // slot range of the row in the CSR

    __device__ __forceinline__ int slot_of(int r) const { return (int)((perm >> (4 * r)) & 15u); }
    __device__ __forceinline__ void clear(int r) {
#if TRB_TC_COUNT_FMA == 2
        ci[r] = 0u;
#elif TRB_TC_COUNT_FMA
        cf[r] = 0ull;
#else
        cnt[r] = 0;
#endif
    }
    __device__ __forceinline__ int count_of(int r) const {
#if TRB_TC_COUNT_FMA == 2
        return (int)(ci[r] / 127u);
#elif TRB_TC_COUNT_FMA
        return __float2int_rn(__uint_as_float((uint32_t)cf[r]) + __uint_as_float((uint32_t)(cf[r] >> 32)));
#else
        return cnt[r];
#endif
    }
    // take n hits back from register slot r (correction paths of TRB_TC_COUNT_FIRST)
    __device__ __forceinline__ void sub(int r, int n) {
#if TRB_TC_COUNT_FMA == 2
        ci[r] -= 127u * (uint32_t)n;
#elif TRB_TC_COUNT_FMA
        cf[r] = (cf[r] & 0xffffffff00000000ull) | (unsigned long long)__float_as_uint(__uint_as_float((uint32_t)cf[r]) - (float)n);
#else
        cnt[r] -= n;
#endif
    }
    // fp32 counters are exact up to 2^24, the scaled integer counters up to 2^32 / 127: the epilogue drains them into the
    // global counters every 256 tiles (<= 16384 hits per lane)
    __device__ __forceinline__ void drain(const Params& p) {
#pragma unroll
        for (int r = 0; r < RTN; ++r) {
            const int c = count_of(r);
            if (s_lo + slot_of(r) < s_hi && c) atomicAdd(p.cnt + s_lo + slot_of(r), c);
            clear(r);
        }
    }
    static __device__ __forceinline__ float enc(float t) {       // register form of a compare value
#if TRB_TC_COUNT_FMA
        return -t * COUNT_SCALE;                                  // power-of-two scaling: exact (+inf -> -inf: never counts)
#else
        return t;
#endif
    }

    __device__ __forceinline__ void init(const Params& p, int64_t q) {
#pragma unroll
        for (int k = 0; k < TRB_TOPK; ++k) { ts[k] = -CUDART_INF_F; tr[k] = -1; }
        float th[RTN];
        int slot[RTN];
#pragma unroll
        for (int r = 0; r < RTN; ++r) { th[r] = CUDART_INF_F; sw[r] = INT32_MAX; clear(r); slot[r] = r; }
        s_lo = s_hi = 0;
        slow = 0;
        if (q >= 0 && p.rel_ptr != nullptr) {
            s_lo = p.rel_ptr[q];
            s_hi = p.rel_ptr[q + 1];
#pragma unroll
            for (int r = 0; r < RTN; ++r)
                if (s_lo + r < s_hi) {
                    const float t = p.thr[s_lo + r];
#if TRB_TC_COUNT_FMA
                    if (!(fabsf(t) >= COUNT_MIN_ABS)) { slow |= 1u << r; continue; }      // also NaN
#endif
                    th[r] = t;
                    const int64_t l = p.thr_gidx[s_lo + r] - p.g_base;
                    // item on a lower shard: the whole local stream lies after it (switch at once); on a higher shard: never
                    sw[r] = l < 0 ? INT32_MIN : (l >= p.G ? INT32_MAX : (int)l - (CH - 1));
                }
        }
        // sort the register slots by threshold, descending (odd-even transposition network, static indices)
#pragma unroll
        for (int pass = 0; pass < RTN; ++pass) {
#pragma unroll
            for (int r = pass & 1; r + 1 < RTN; r += 2) {
                if (th[r] < th[r + 1]) {
                    const float tf = th[r]; th[r] = th[r + 1]; th[r + 1] = tf;
                    const int ti = sw[r]; sw[r] = sw[r + 1]; sw[r + 1] = ti;
                    const int tp = slot[r]; slot[r] = slot[r + 1]; slot[r + 1] = tp;
                }
            }
        }
        perm = 0;
        next_sw = INT32_MAX;
#pragma unroll
        for (int r = 0; r < RTN; ++r) {
            perm |= (uint32_t)slot[r] << (4 * r);
            te[r] = th[r] == CUDART_INF_F ? enc(th[r]) : enc(next_below(th[r]));     // unused slot: never exceeded
            next_sw = min(next_sw, sw[r]);
        }
    }

    // The stream of this warp has reached chunk g0 >= next_sw: every item whose row is < g0 + 32 now compares strictly
    // (te = thr).  Returns whether one of them lies INSIDE the chunk (the caller then corrects for ties that precede it).
    __device__ __forceinline__ bool pass_items(const Params& p, int g0) {
        bool own = false;
        int nxt = INT32_MAX;
#pragma unroll
        for (int r = 0; r < RTN; ++r) {
            if (g0 >= sw[r]) {
                te[r] = enc(p.thr[s_lo + slot_of(r)]);
                own |= sw[r] != INT32_MIN && g0 <= sw[r] + (CH - 1);
                sw[r] = INT32_MAX;
            }
            nxt = min(nxt, sw[r]);
        }
        next_sw = nxt;
        return own;
    }

    // the stream is in ascending index order, so a later value only displaces strictly smaller entries
    __device__ __forceinline__ void insert(float s, int grow) {
        if (!(s > ts[TRB_TOPK - 1])) return;
        bool placed = false;
#pragma unroll
        for (int k = TRB_TOPK - 1; k >= 1; --k) {
            if (!placed) {
                if (s > ts[k - 1]) { ts[k] = ts[k - 1]; tr[k] = tr[k - 1]; }
                else { ts[k] = s; tr[k] = grow; placed = true; }
            }
        }
        if (!placed) { ts[0] = s; tr[0] = grow; }
    }

    __device__ __forceinline__ void flush(const Params& p, int64_t q, int64_t nlists, int64_t list) {
        float* cs = p.cand_sim + (q * nlists + list) * TRB_TOPK;
        int64_t* ci = p.cand_idx + (q * nlists + list) * TRB_TOPK;
#pragma unroll
        for (int k = 0; k < TRB_TOPK; ++k) { cs[k] = ts[k]; ci[k] = tr[k] >= 0 ? p.g_base + tr[k] : INT64_MAX; }
        drain(p);
    }
};

__device__ __forceinline__ void pad_list(const Params& p, int64_t q, int64_t nlists, int64_t list) {
    float* cs = p.cand_sim + (q * nlists + list) * TRB_TOPK;
    int64_t* ci = p.cand_idx + (q * nlists + list) * TRB_TOPK;
#pragma unroll
    for (int k = 0; k < TRB_TOPK; ++k) { cs[k] = -CUDART_INF_F; ci[k] = INT64_MAX; }
}

// Cold paths, working from a local copy of the chunk and the thresholds in global memory; corrections go straight to
// the global counters.
//  fix_own_chunk : the chunk contains relevant items of this row; the fast path compared it against thr itself (">"),
//                  values equal to thr that precede the item must be added.
//  count_overflow: rows with more than RTN relevant items -- exact count of the extra slots for this chunk.
#if !TRB_TC_COUNT_FIRST
__device__ __noinline__ void fix_own_chunk(const Params& p, const float* lv, int g0, int64_t s_lo, int64_t s_end, uint32_t slow) {
    for (int64_t slot = s_lo; slot < s_end; ++slot) {
        if ((slow >> (int)(slot - s_lo)) & 1u) continue;           // counted exactly by count_exact
        const int64_t lg = p.thr_gidx[slot] - p.g_base;            // local row of the item
        // an item of a HIGHER shard whose index falls into this shard's zero-padded tail is not in this chunk: its slot
        // never switched to the strict compare, so the ties before it are already counted
        if (lg >= p.G) continue;
        const int64_t l = lg - g0;                                 // position of the item inside this chunk
        if (l < 0 || l >= CH) continue;
        const float th = p.thr[slot];
        int c = 0;
        for (int j = 0; j < (int)l; ++j) c += (lv[j] == th) ? 1 : 0;
        if (c) atomicAdd(p.cnt + slot, c);
    }
}
#endif
// TRB_TC_COUNT_FIRST: the chunk was counted against nextbelow(thr) (">=" semantics) although the slot switches to the strict
// compare here: values equal to thr AT or AFTER the item (position l in the chunk; l < 0: the item precedes the chunk) go back.
__device__ __noinline__ void unfix_switched_slot(const Params& p, const float* lv, int g0, int64_t slot) {
    const int64_t l = p.thr_gidx[slot] - p.g_base - g0;
    const float th = p.thr[slot];
    int c = 0;
    for (int j = l > 0 ? (int)l : 0; j < CH; ++j) c += (lv[j] == th) ? 1 : 0;
    if (c) atomicAdd(p.cnt + slot, -c);
}
// exact count of this chunk for the CSR positions >= rtn (rows with more relevant items than register slots) and the
// positions flagged in `slow`
__device__ __noinline__ void count_exact(const Params& p, const float* lv, int g0, int64_t s_lo, int64_t s_hi, int rtn, uint32_t slow) {
    for (int64_t slot = s_lo; slot < s_hi; ++slot) {
        const int i = (int)(slot - s_lo);
        if (i < rtn && !((slow >> i) & 1u)) continue;
        const float th = p.thr[slot];
        const int64_t ti = p.thr_gidx[slot] - p.g_base - g0;
        int c = 0;
        for (int j = 0; j < CH; ++j) c += (lv[j] > th || (lv[j] == th && j < ti)) ? 1 : 0;
        if (c) atomicAdd(p.cnt + slot, c);
    }
}

// the register-resident rank counts of one chunk: every value against every threshold slot of the row
template <int RTN>
__device__ __forceinline__ void count_chunk(RowState<RTN>& st, const float (&v)[CH], float cmax) {
#pragma unroll
    for (int r = 0; r < RTN; ++r) {
        // no value of the chunk, in any row of the warp, reaches the r-th best threshold: nothing to count (thresholds of
        // relevant items sit in the upper tail of the similarity distribution, so the first slots skip most chunks)
#if TRB_TC_COUNT_FMA
        const float c = st.te[r];                  // = -te * 2^100
        if (r < TRB_TC_SKIP_R && !__any_sync(0xffffffffu, fmaf(cmax, COUNT_SCALE, c) > 0.f)) continue;
#if TRB_TC_COUNT_FMA == 2
        uint32_t acc = st.ci[r];
#pragma unroll
        for (int j = 0; j < CH; j += 3) {
            uint32_t a = __float_as_uint(__saturatef(fmaf(v[j], COUNT_SCALE, c)));
            if (j + 1 < CH) a += __float_as_uint(__saturatef(fmaf(v[j + 1], COUNT_SCALE, c)));
            if (j + 2 < CH) a += __float_as_uint(__saturatef(fmaf(v[j + 2], COUNT_SCALE, c)));
            acc += a >> 23;
        }
        st.ci[r] = acc;
#else
        unsigned long long acc = st.cf[r];
#pragma unroll
        for (int j = 0; j < CH; j += 2)            // 2 x FFMA.SAT (immediate form) + 1 x FADD2: 1.5 instructions per value, FMA pipe
            add2(acc, __saturatef(fmaf(v[j], COUNT_SCALE, c)), __saturatef(fmaf(v[j + 1], COUNT_SCALE, c)));
        st.cf[r] = acc;
#endif
#else
        if (r < TRB_TC_SKIP_R && !__any_sync(0xffffffffu, cmax > st.te[r])) continue;
        const float te = st.te[r];
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
            c0 += __float_as_uint(te - v[j]) >> 31;           // te - v < 0  <=>  v > te   (x - x = +0, never -0)
            c1 += __float_as_uint(te - v[j + 1]) >> 31;
            c2 += __float_as_uint(te - v[j + 2]) >> 31;
            c3 += __float_as_uint(te - v[j + 3]) >> 31;
        }
        st.cnt[r] += (int)((c0 + c1) + (c2 + c3));
#endif
    }
}

// One chunk (CH accumulator columns of a row, in registers) through the epilogue.  The common chunk -- no top-10 candidate, no
// threshold switching to its strict compare, no row that needs the exact slow path, no padded tail tile, in ANY lane of the warp --
// costs the chunk maximum, one warp-uniform vote and the counting loop.  With TRB_TC_COUNT_FIRST (shipped) the counting loop runs
// BEFORE the vote is consumed, so that the vote's dependent instructions resolve under the arithmetic, and the rare path corrects
// the counts afterwards; otherwise the rare path prepares (switches thresholds, masks the tail) and counts itself.
// chunk maximum (and the maxima of its 4-value groups) for the top-10 filter; ptxas folds this into 3-input FMNMX3
__device__ __forceinline__ float chunk_max(const float (&v)[CH], float (&m8)[CH / 4]) {
    constexpr int NG = CH / 4;
#pragma unroll
    for (int i = 0; i < NG; ++i) m8[i] = fmaxf(fmaxf(v[4 * i], v[4 * i + 1]), fmaxf(v[4 * i + 2], v[4 * i + 3]));
    float cmax = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
    if (NG == 8) cmax = fmaxf(cmax, fmaxf(fmaxf(m8[NG - 4], m8[NG - 3]), fmaxf(m8[NG - 2], m8[NG - 1])));
    return cmax;
}

// `tail` (a register flag, the same for every chunk of a tile): the tile holds zero-padded gallery rows, which must not rank.
// Returns (warp-uniformly) whether the rare path ran, i.e. whether the lanes may have diverged.
template <int RTN>
__device__ __forceinline__ bool stream_chunk(const Params& p, RowState<RTN>& st, float (&v)[CH], int g0, bool row_valid,
                                             bool warp_has_thr, uint32_t tail, bool row_slow) {
    constexpr int NG = CH / 4;
    float m8[NG];
    float cmax = chunk_max(v, m8);
#if TRB_TC_COUNT_FIRST
    static_assert(TRB_TC_COUNT_FMA != 0, "count-first needs the scaled compare values");
    {
#ifdef TRB_TC_PROBE
        if (p.debug & 8) row_valid = false;
#endif
        const bool rare = (row_valid && cmax > st.ts[TRB_TOPK - 1]) || g0 >= st.next_sw || row_slow || tail != 0u;
        const bool any_rare = __any_sync(0xffffffffu, rare);
        if (warp_has_thr) count_chunk<RTN>(st, v, cmax);       // unconditional: the vote resolves under these instructions
        if (!any_rare) return false;
    }
    if (tail != 0u) {                          // padded gallery rows (zero vectors, similarity exactly 0) never rank
        const int gvalid = (int)p.G;
        const int npad = min(max(g0 + CH - gvalid, 0), CH);
        if (warp_has_thr && npad > 0) {
#pragma unroll
            for (int r = 0; r < RTN; ++r)
                if (st.te[r] > 0.f) st.sub(r, npad);           // te holds -t * 2^100: the zeros were counted iff t < 0
        }
#pragma unroll
        for (int j = 0; j < CH; ++j)
            if (g0 + j >= gvalid) v[j] = -CUDART_INF_F;
        cmax = chunk_max(v, m8);
    }
#else
#if TRB_TC_FAST_CHUNK
    {
#ifdef TRB_TC_PROBE
        if (p.debug & 8) row_valid = false;
#endif
        // the tail flag rides in the same test: no separate per-chunk branch for the one padded tile of a gallery
        const bool rare = (row_valid && cmax > st.ts[TRB_TOPK - 1]) || g0 >= st.next_sw || row_slow || tail != 0u;
        if (!__any_sync(0xffffffffu, rare)) {
            if (warp_has_thr) count_chunk<RTN>(st, v, cmax);
            return false;
        }
    }
#endif
    if (tail != 0u) {                          // padded gallery rows (zero vectors) never rank
        const int gvalid = (int)p.G;
#pragma unroll
        for (int j = 0; j < CH; ++j)
            if (g0 + j >= gvalid) v[j] = -CUDART_INF_F;
        cmax = chunk_max(v, m8);
    }

#endif
    // ---- top-10: candidates are rare after the first tiles; only 4-value groups whose maximum beats the current
    //      10th best are scanned, and the hits go through a bit mask + select tree (keeps the hot loop compact) ----
#ifdef TRB_TC_PROBE
    if (p.debug & 8) row_valid = false;        // probe: rank counts only, no top-10 maintenance
#endif
    if (row_valid && cmax > st.ts[TRB_TOPK - 1]) {
        const float kth = st.ts[TRB_TOPK - 1];
        uint32_t cand = 0;
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            if (m8[i] > kth) {
                uint32_t m = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) m |= (v[4 * i + j] > kth) ? (1u << j) : 0u;
                cand |= m << (4 * i);
            }
        }
#pragma unroll 1
        while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1;
            st.insert(select_lane(v, j), g0 + j);
        }
    }

    // ---- exact rank counts ----
    if (!warp_has_thr) return true;
#if TRB_TC_COUNT_FIRST
    const bool switching = row_valid && g0 >= st.next_sw;
    if (switching || row_slow) {
        float lv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) lv[j] = v[j];
        if (switching) {
#pragma unroll
            for (int r = 0; r < RTN; ++r)
                if (g0 >= st.sw[r]) unfix_switched_slot(p, lv, g0, st.s_lo + st.slot_of(r));
        }
        if (row_slow) count_exact(p, lv, g0, st.s_lo, st.s_hi, RTN, st.slow);
    }
    if (g0 >= st.next_sw) st.pass_items(p, g0);            // the following chunks compare strictly
#else
    bool own = false;
    if (g0 >= st.next_sw) own = st.pass_items(p, g0);      // rare: 1 + (relevant items of the row) times per stream
    count_chunk<RTN>(st, v, cmax);
    const bool overflow = row_slow;
    own = own && row_valid;
    if (own || overflow) {
        float lv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) lv[j] = v[j];
        if (own) fix_own_chunk(p, lv, g0, st.s_lo, min(st.s_lo + RTN, st.s_hi), st.slow);
        if (overflow) count_exact(p, lv, g0, st.s_lo, st.s_hi, RTN, st.slow);
    }
#endif
    return true;
}

// COUNTS: the launch carries thresholds (rank counts wanted).  A compile-time constant so that the chunk loop has no "does this warp
// have thresholds" branch; a warp whose 32 rows happen to have none counts against -inf (nothing), which is rare and harmless.
template <int MODE, int RTN, bool COUNTS>
__global__ void __launch_bounds__(NUM_THREADS, 1) retrieval_tc_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KC = p.kchunks, NS = p.nstages;
    uint8_t* sA = smem;                                   // KC x 16 KiB: the query tile, all of K
    uint8_t* sB = sA + (size_t)KC * BLOCK_BYTES;          // NS x 32 KiB ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)NS * STAGE_BYTES);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* t_full = bars + 2;            // [NTB]
    uint64_t* t_empty = bars + 2 + NTB;     // [NTB]
    uint64_t* b_full = bars + 2 + 2 * NTB;  // [NS]
    uint64_t* b_empty = bars + 2 + 2 * NTB + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * NTB + 2 * MAX_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < NTB; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, NUM_EPI_WARPS); }
        for (int i = 0; i < NS; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------- producer -------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t bphase = 0, aphase = 0;
            for (int64_t u = blockIdx.x; u < p.num_units; u += gridDim.x) {
                const UnitInfo ui = unit_info<MODE>(p, u);
                if (ui.t_lo >= ui.t_hi) continue;
                role_wait(a_empty, aphase ^ 1, 256);
                mbar_expect_tx(a_full, (uint32_t)KC * BLOCK_BYTES);
                for (int kc = 0; kc < KC; ++kc)
                    bulk_g2s(sA + (size_t)kc * BLOCK_BYTES, p.q_packed + ((size_t)ui.qt * KC + kc) * BLOCK_BYTES, BLOCK_BYTES, a_full);
                aphase ^= 1;
                for (int64_t t = ui.t_lo; t < ui.t_hi; ++t) {
                    for (int kc = 0; kc < KC; ++kc) {
                        role_wait(b_empty + stage, bphase ^ 1, 128);
#ifdef TRB_TC_PROBE
                        // probe: after the ring has been filled once, only signal (no L2 traffic; the MMAs re-use stale tiles)
                        if ((p.debug & 4) && (t - ui.t_lo) * KC + kc >= NS) {
                            mbar_arrive(b_full + stage);
                            if (++stage == NS) { stage = 0; bphase ^= 1; }
                            continue;
                        }
#endif
                        mbar_expect_tx(b_full + stage, STAGE_BYTES);
                        uint8_t* dst = sB + (size_t)stage * STAGE_BYTES;
#pragma unroll
                        for (int b = 0; b < TILE_BLOCKS; ++b)
                            bulk_g2s(dst + b * BLOCK_BYTES, p.g_packed + ((size_t)(TILE_BLOCKS * t + b) * KC + kc) * BLOCK_BYTES, BLOCK_BYTES,
                                     b_full + stage);
                        if (++stage == NS) { stage = 0; bphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer -----------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, TILE_N);
            int stage = 0, tbuf = 0;
            uint32_t bphase = 0, aphase = 0, tphase = 0;
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
            for (int64_t u = blockIdx.x; u < p.num_units; u += gridDim.x) {
                const UnitInfo ui = unit_info<MODE>(p, u);
                if (ui.t_lo >= ui.t_hi) continue;
                role_wait(a_full, aphase, 128);
                aphase ^= 1;
                tc_fence_after();
                for (int64_t t = ui.t_lo; t < ui.t_hi; ++t) {
                    TRB_STAMP(14, t);
                    role_wait(t_empty + tbuf, tphase ^ 1, 256);     // epilogue has drained this accumulator
                    tc_fence_after();
                    TRB_STAMP(0, t);
                    const uint32_t d_tmem = tmem_base + (uint32_t)tbuf * TILE_N;
                    for (int kc = 0; kc < KC; ++kc) {
                        role_wait(b_full + stage, bphase, 64);
                        tc_fence_after();
                        const uint32_t a0 = a_addr + (uint32_t)kc * BLOCK_BYTES;
                        const uint32_t b0 = b_addr + (uint32_t)stage * STAGE_BYTES;
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                            umma_bf16(d_tmem, umma_desc_sw128(a0 + kk * UMMA_K * 2), umma_desc_sw128(b0 + kk * UMMA_K * 2), idesc,
                                      (uint32_t)((kc | kk) != 0));
                        umma_commit(b_empty + stage);          // smem slot reusable once these MMAs retire
                        if (++stage == NS) { stage = 0; bphase ^= 1; }
                    }
                    umma_commit(t_full + tbuf);                // accumulator complete -> epilogue
                    TRB_STAMP(1, t);
                    if (++tbuf == NTB) { tbuf = 0; tphase ^= 1; }
                }
                umma_commit(a_empty);                          // query tile may be overwritten
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ------------------------------- epilogue -------------------------------------------
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int colgrp = (warp - EPI_WARP0) >> 2;      // which share of the 256 accumulator columns (chunks_of / chunk_col)
        const int row = quarter * 32 + lane;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int my_chunks = chunks_of(colgrp);
        int tbuf = 0;
        uint32_t tphase = 0;
        for (int64_t u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const UnitInfo ui = unit_info<MODE>(p, u);
            if (ui.t_lo >= ui.t_hi) continue;
            const int64_t prow = ui.qt * TILE_M + row;                 // packed query row
            const int64_t q = prow < p.Q ? p.q_row_id[prow] : -1;      // original query number
            RowState<RTN> st;
            int32_t blo = 0, bhi = 0;
            int64_t slot_base = 0;
            if (MODE == 0) {
                st.init(p, q);
            } else if (q >= 0) {
                blo = p.band_lo[prow];
                bhi = p.band_hi[prow];
                slot_base = p.rel_ptr[q] + p.rel_off[prow];
            }
            constexpr bool warp_has_thr = MODE == 0 && COUNTS;
            // rows with more relevant items than register slots, or with a threshold too close to zero for the scaled compare
            const bool row_slow = MODE == 0 && q >= 0 && ((st.s_hi - st.s_lo > RTN) || st.slow != 0u);

            // Only the last tile of the gallery holds zero padding rows.  The tile body exists twice, with the tail handling as a
            // compile-time constant: the hot copy carries no per-chunk test against p.G (the compiler re-derived a run-time flag
            // inside every chunk: a constant-bank load and a 64-bit compare, 5 % of the chunk loop), the cold copy runs once per unit.
            // (RTN = 8: one copy with a run-time flag -- two copies of its 6 KB chunk loop cost more in instruction fetch, 81 vs 72 ms.)
            auto tile_body = [&](const int64_t t, auto tail_c) {
                const uint32_t tail_flag = tail_value(tail_c);
                mbar_wait_sleepy(t_full + tbuf, tphase, p.wait_hint_ns);
                tc_fence_after();
                if (lane == 0 && (warp == 2 || warp == 6)) TRB_STAMP(warp == 2 ? 2 : 4, t);
                const uint32_t taddr0 = lane_taddr + (uint32_t)(tbuf * TILE_N);
                auto consume = [&](float (&v)[CH], int chunk) {
                    const int g0 = (int)(t * TILE_N) + chunk_col(colgrp, chunk);   // packed gallery row of v[0]
#ifdef TRB_TC_PROBE
                    if (p.debug & 3) { if (v[5] == 12345.678f) p.cand_sim[0] = v[7]; return; }
#endif
                    if (MODE == 1) {
                        if (q >= 0 && g0 < bhi && g0 + CH > blo) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) {
                                const int g = g0 + j;
                                if (g >= blo && g < bhi) {
                                    p.thr[slot_base + (g - blo)] = v[j];
                                    p.thr_gidx[slot_base + (g - blo)] = p.g_row_id[g];
                                }
                            }
                        }
                        return;
                    }
                    stream_chunk<RTN>(p, st, v, g0, q >= 0, warp_has_thr, tail_flag, row_slow);
                };
                auto release = [&]() {                     // accumulator fully read by this warp: hand it back
                    tc_fence_before();
                    if (lane == 0) TRB_STAMP(6 + warp - EPI_WARP0, t);
                    if (lane == 0) mbar_arrive(t_empty + tbuf);
                };
#if TRB_TC_PREFETCH
                // two register buffers: the TMEM load of chunk c+1 is in flight while chunk c is consumed
                static_assert(COLGRP_CONTIGUOUS && (TILE_CHUNKS / NUM_COLGRP) % 2 == 0, "the prefetching epilogue consumes chunks in pairs");
                constexpr int CHUNKS_PER_WARP = TILE_CHUNKS / NUM_COLGRP;
                float va[CH], vb[CH];
                __syncwarp();                              // tcgen05.ld is warp-collective (.sync.aligned)
                tmem_ld_issue(taddr0 + (uint32_t)chunk_col(colgrp, 0), va);
                tmem_ld_wait(va);
#pragma unroll 1
                for (int chunk = 0; chunk < CHUNKS_PER_WARP; chunk += 2) {
                    __syncwarp();
                    tmem_ld_issue(taddr0 + (uint32_t)chunk_col(colgrp, chunk + 1), vb);
                    consume(va, chunk);
                    tmem_ld_wait(vb);
                    const bool more = chunk + 2 < CHUNKS_PER_WARP;
                    __syncwarp();
                    if (more) tmem_ld_issue(taddr0 + (uint32_t)chunk_col(colgrp, chunk + 2), va);
                    else release();
                    consume(vb, chunk + 1);
                    if (more) tmem_ld_wait(va);
                }
#else
#pragma unroll 1                           // a fully unrolled loop (17 KB) falls out of the instruction cache: 85 vs 57 ms
                for (int chunk = 0; chunk < my_chunks; ++chunk) {
                    float v[CH];
                    __syncwarp();                          // tcgen05.ld is warp-collective (.sync.aligned)
                    tmem_ld_issue(taddr0 + (uint32_t)chunk_col(colgrp, chunk), v);
                    tmem_ld_wait(v);
                    if (chunk == my_chunks - 1) release();
                    consume(v, chunk);
                }
#endif
                if (lane == 0 && (warp == 2 || warp == 6)) TRB_STAMP(warp == 2 ? 3 : 5, t);

                if (++tbuf == NTB) { tbuf = 0; tphase ^= 1; }
                if (MODE == 0 && TRB_TC_COUNT_FMA && ((t - ui.t_lo) & 255) == 255 && warp_has_thr) st.drain(p);
            };
            if (RTN <= 4) {
                const bool has_tail = MODE == 0 && ui.t_hi * TILE_N > p.G;      // then it is tile t_hi - 1
                const int64_t t_fast_end = has_tail ? ui.t_hi - 1 : ui.t_hi;
                for (int64_t t = ui.t_lo; t < t_fast_end; ++t) tile_body(t, std::false_type{});
                if (has_tail) tile_body(ui.t_hi - 1, std::true_type{});
            } else {
                for (int64_t t = ui.t_lo; t < ui.t_hi; ++t) tile_body(t, (uint32_t)((t + 1) * TILE_N > p.G ? 1u : 0u));
            }

            if (MODE == 0 && q >= 0) {
                const int64_t nlists = LISTS_PER_SPLIT * (int64_t)p.nsplit;
                st.flush(p, q, nlists, (ui.split < 0 ? 0 : ui.split) * LISTS_PER_SPLIT + colgrp);
                if (ui.split < 0)            // unsplit unit: the list groups of the other splits stay empty
                    for (int64_t l = LISTS_PER_SPLIT + colgrp; l < nlists; l += LISTS_PER_SPLIT)
                        pad_list(p, q, nlists, l);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// pack: optional gather + L2 normalise + round to bf16 + write the pre-swizzled tile-major image
// one warp per packed row; each lane moves 16-byte (8 x bf16) chunks
// ---------------------------------------------------------------------------------------------
template <bool SRC_BF16>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const void* __restrict__ src, const int64_t* __restrict__ perm, int normalize, float eps,
                 uint8_t* __restrict__ packed, int64_t rows, int64_t rows_padded, int64_t dim) {
    const int lane = threadIdx.x & 31;
    const int64_t prow = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (prow >= rows_padded) return;
    const int64_t nchunk = dim >> 3, kchunks = dim >> 6;
    if (prow >= rows) {
        for (int64_t c = lane; c < nchunk; c += 32)
            *reinterpret_cast<uint4*>(packed + packed_offset_bytes(prow, c, kchunks)) = make_uint4(0, 0, 0, 0);
        return;
    }
    const int64_t srow = perm ? perm[prow] : prow;
    float ss = 0.f;
    // first pass: norm (skipped when not normalising)
    if (normalize) {
        for (int64_t c = lane; c < nchunk; c += 32) {
            float x[8];
            if (SRC_BF16) {
                const uint4 raw = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(src) + srow * dim + c * 8);
                const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __bfloat162float(h[i]);
            } else {
                const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(src) + srow * dim + c * 8);
                const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(src) + srow * dim + c * 8 + 4);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(x[i], x[i], ss);
        }
        ss = warp_sum(ss);
    }
    const float nrm = normalize ? fmaxf(sqrtf(ss), eps) : 1.0f;
    for (int64_t c = lane; c < nchunk; c += 32) {
        float x[8];
        if (SRC_BF16) {
            const uint4 raw = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(src) + srow * dim + c * 8);
            const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __bfloat162float(h[i]);
        } else {
            const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(src) + srow * dim + c * 8);
            const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(src) + srow * dim + c * 8 + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        }
        uint4 out;
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(&out);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __float2bfloat16_rn(normalize ? __fdiv_rn(x[i], nrm) : x[i]);
        *reinterpret_cast<uint4*>(packed + packed_offset_bytes(prow, c, kchunks)) = out;
    }
}

}  // namespace

extern "C" int trb_retrieval_tc_lists_per_split(void) { return LISTS_PER_SPLIT; }

extern "C" int64_t trb_packed_rows(int64_t rows) { return rows <= 0 ? 0 : ((rows + PACK_ROWS - 1) / PACK_ROWS) * PACK_ROWS; }

extern "C" int64_t trb_packed_bytes(int64_t rows, int64_t dim) {
    if (dim <= 0 || dim % 64 != 0) return 0;
    return trb_packed_rows(rows) * dim * 2;
}

extern "C" int trb_pack_rows_bf16(const void* src, int src_is_bf16, const int64_t* perm, int normalize, float eps, void* packed,
                                  int64_t rows, int64_t dim, trb_stream_t stream) {
    TRB_REQUIRE(src && packed, "pack_rows: null pointer");
    TRB_REQUIRE(rows >= 0 && dim > 0, "pack_rows: bad shape");
    if (dim % 64 != 0) { trb_set_error("pack_rows: dim=%lld must be a multiple of 64 on the tensor-core path", (long long)dim); return TRB_ERR_UNSUPPORTED; }
    TRB_REQUIRE(trb_aligned16(src) && trb_aligned16(packed), "pack_rows: buffers must be 16-byte aligned");
    const int64_t rp = trb_packed_rows(rows);
    if (rp == 0) return 0;
    const unsigned grid = (unsigned)trb_ceil_div(rp, 8);
    if (src_is_bf16)
        pack_rows_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, perm, normalize, eps, (uint8_t*)packed, rows, rp, dim);
    else
        pack_rows_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(src, perm, normalize, eps, (uint8_t*)packed, rows, rp, dim);
    TRB_LAUNCH_OK();
    return 0;
}

extern "C" int trb_retrieval_stream_tc(const void* q_packed, const void* g_packed, int64_t Q, int64_t G, int64_t D,
                                       const int64_t* q_row_id, const int64_t* g_row_id, int64_t g_base, const int64_t* rel_ptr, float* thr,
                                       int64_t* thr_gidx, const int32_t* band_lo, const int32_t* band_hi, const int32_t* rel_off,
                                       int mode, int nsplit, int max_rel, float* cand_sim, int64_t* cand_idx, int32_t* cnt,
                                       trb_stream_t stream) {
    TRB_REQUIRE(q_packed && g_packed && q_row_id, "stream_tc: null pointer");
    TRB_REQUIRE(mode == 0 || g_row_id != nullptr, "stream_tc: mode 1 needs g_row_id (global index of the pid-sorted packed rows)");
    TRB_REQUIRE(Q >= 0 && G >= 0 && G < (1LL << 31) - 512, "stream_tc: bad shape (a shard holds < 2^31 gallery rows)");
    TRB_REQUIRE(mode == 0 || mode == 1, "stream_tc: mode must be 0 (stream) or 1 (threshold capture)");
    if (D <= 0 || D % 64 != 0 || D > 512) {
        trb_set_error("stream_tc: D=%lld unsupported on the tensor-core path (needs a multiple of 64, <= 512)", (long long)D);
        return TRB_ERR_UNSUPPORTED;
    }
    TRB_REQUIRE((reinterpret_cast<uintptr_t>(q_packed) & 127) == 0 && (reinterpret_cast<uintptr_t>(g_packed) & 127) == 0,
                "stream_tc: packed operands must be 128-byte aligned");
    if (mode == 0) {
        TRB_REQUIRE(cand_sim && cand_idx, "stream_tc: candidate buffers are required in mode 0");
        TRB_REQUIRE(nsplit >= 1, "stream_tc: nsplit must be >= 1");
        TRB_REQUIRE((rel_ptr == nullptr) == (thr == nullptr) && (thr == nullptr) == (thr_gidx == nullptr) &&
                        (thr == nullptr) == (cnt == nullptr),
                    "stream_tc: rel_ptr, thr, thr_gidx and cnt must be given together");
    } else {
        TRB_REQUIRE(rel_ptr && thr && thr_gidx && band_lo && band_hi && rel_off, "stream_tc: mode 1 needs rel_ptr, thr, thr_gidx and the band arrays");
        nsplit = 1;
    }
    if (Q == 0 || G == 0) return 0;

    Params p;
    p.q_packed = static_cast<const uint8_t*>(q_packed);
    p.g_packed = static_cast<const uint8_t*>(g_packed);
    p.Q = Q; p.G = G;
    p.kchunks = (int)(D / 64);
    const int a_bytes = p.kchunks * BLOCK_BYTES;
    const int fixed_bytes = (2 + 2 * NTB + 2 * MAX_STAGES) * 8 + 16 + 1024;   // barriers + TMEM slot + alignment slack
    int ns = (SMEM_MAX - fixed_bytes - a_bytes) / STAGE_BYTES;
    p.nstages = ns > MAX_STAGES ? MAX_STAGES : ns;
    TRB_REQUIRE(p.nstages >= 2, "stream_tc: not enough shared memory for a 2-stage ring at D=%lld", (long long)D);
    p.q_row_id = q_row_id; p.g_row_id = g_row_id; p.g_base = g_base; p.rel_ptr = rel_ptr; p.thr = thr; p.thr_gidx = thr_gidx;
    p.band_lo = band_lo; p.band_hi = band_hi; p.rel_off = rel_off;
    p.nsplit = nsplit; p.cand_sim = cand_sim; p.cand_idx = cand_idx; p.cnt = cnt;
    p.wait_hint_ns = getenv("TRB_TC_WAIT_NS") ? (uint32_t)atoi(getenv("TRB_TC_WAIT_NS")) : 1000u;
    p.debug = getenv("TRB_TC_DEBUG") ? atoi(getenv("TRB_TC_DEBUG")) : 0;
    p.num_qtiles = trb_ceil_div(Q, TILE_M);
    p.num_gtiles = trb_ceil_div(G, TILE_N);
    TRB_REQUIRE(nsplit <= p.num_gtiles, "stream_tc: nsplit=%d exceeds the number of gallery tiles %lld", nsplit, (long long)p.num_gtiles);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // whole waves of the persistent grid take unsplit query tiles; only the remainder is split (nsplit pieces each)
    p.full_qtiles = (mode == 0 && nsplit > 1) ? (p.num_qtiles / sms) * sms : (mode == 0 ? p.num_qtiles : 0);
    // (splitting EVERY query tile so that a gallery piece stays L2-resident was measured and loses: 64 / 67 / 74 ms at 4 / 8 / 16
    //  pieces against 57 ms -- the per-unit start-up outweighs the DRAM re-reads, which run at < 2 % of the HBM bandwidth)
    p.num_units = mode == 0 ? p.full_qtiles + (p.num_qtiles - p.full_qtiles) * nsplit : p.num_qtiles;

    const int smem_bytes = a_bytes + p.nstages * STAGE_BYTES + fixed_bytes;
    const unsigned grid = (unsigned)(p.num_units < sms ? p.num_units : sms);
#define TRB_LAUNCH_TC(MODE_, RTN_, COUNTS_)                                                                                      \
    do {                                                                                                                         \
        TRB_CUDA_OK(cudaFuncSetAttribute(retrieval_tc_kernel<MODE_, RTN_, COUNTS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         smem_bytes));                                                                           \
        retrieval_tc_kernel<MODE_, RTN_, COUNTS_><<<grid, NUM_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);                    \
    } while (0)
    if (mode == 1) TRB_LAUNCH_TC(1, 4, false);
    else if (rel_ptr == nullptr) TRB_LAUNCH_TC(0, 4, false);      // top-k only
    else if (max_rel <= 4) TRB_LAUNCH_TC(0, 4, true);
    else TRB_LAUNCH_TC(0, 8, true);
#undef TRB_LAUNCH_TC
#ifdef TRB_TC_PROBE
    if ((p.debug & 16) && mode == 0) {
        unsigned long long h[16][STAMP_TILES];
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpyFromSymbol(h, g_stamps, sizeof(h));
        fprintf(stderr, "STAMPS tile: mma_start mma_commit | epi2_ready epi2_done | epi6_ready epi6_done   (ns after the first stamp)\n");
        for (int i = 0; i < STAMP_TILES; ++i)
        {
            unsigned long long rmax = 0;
            for (int w = 0; w < 8; ++w) rmax = h[6 + w][i] > rmax ? h[6 + w][i] : rmax;
            fprintf(stderr, "STAMPS %4d: wait %7lld start %7lld commit %7lld | epi2 %7lld..%7lld epi6 %7lld..%7lld | released by all at %7lld; w2..w9:", STAMP_FIRST + i,
                    (long long)(h[14][i] - h[0][0]), (long long)(h[0][i] - h[0][0]), (long long)(h[1][i] - h[0][0]), (long long)(h[2][i] - h[0][0]),
                    (long long)(h[3][i] - h[0][0]), (long long)(h[4][i] - h[0][0]), (long long)(h[5][i] - h[0][0]), (long long)(rmax - h[0][0]));
            for (int w = 0; w < 8; ++w) fprintf(stderr, " %lld", (long long)(h[6 + w][i] - h[0][0]));
            fprintf(stderr, "\n");
        }
    }
#endif
    TRB_LAUNCH_OK();
    return 0;
}
