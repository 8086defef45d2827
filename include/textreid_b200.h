/*
 * libtextreid_b200 -- C ABI of the B200-native TextReID hot path.
 *
 * The reference (BrandonHanx/TextReID) is pure PyTorch and exposes no FFI; its "operator
 * surface" is a set of Python call shapes.  Each entry point below names the reference code
 * it replaces (paths relative to the reference root).  The Python host in `textreid_b200/`
 * binds these with ctypes and mirrors the reference signatures; INTEGRATION.md shows the stub
 * a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named `host_*`; the caller owns all memory,
 *     including outputs and workspaces (sizes from the trb_*_bytes helpers);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise,
 *     never allocate, keep no state between calls, and are CUDA-graph capturable;
 *   - matrices are dense row-major; fp32 unless stated; ids / indices are int64 like the
 *     reference's LongTensors;
 *   - return value: 0 on success, TRB_ERR_* (negative) for rejected arguments, or a positive
 *     cudaError_t.  trb_last_error_string() describes the last failure on the calling thread;
 *   - there is no CPU fallback anywhere in this library.
 *
 * Ranking order (north star): similarity descending, ties by ascending gallery index, i.e.
 * torch.argsort(descending=True, stable=True).  TRB_TOPK = 10 = max(topk) of the reference
 * (lib/engine/inference.py:95 hard-codes topk=[1,5,10]).
 */
#ifndef TEXTREID_B200_H
#define TEXTREID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRB_VERSION 100          /* 0.1.0 */
#define TRB_TOPK_DEPTH 10

#define TRB_ERR_INVALID (-1)     /* bad argument (shape, alignment, null pointer) */
#define TRB_ERR_UNSUPPORTED (-2) /* valid request this build cannot serve (e.g. D not a multiple of 64 on the tensor-core path) */
#define TRB_ERR_WORKSPACE (-3)   /* workspace too small */

typedef void* trb_stream_t;

int trb_version(void);
const char* trb_last_error_string(void);

/* ------------------------------------------------------------------------------------------
 * Row utilities
 * ---------------------------------------------------------------------------------------- */

/* y[r,:] = x[r,:] / max(||x[r,:]||_2, eps); optionally inv_norm[r] = 1/max(||x||, eps).
 * Replaces F.normalize(p=2, dim=1): head.py:128-129,139,145; losses.py:112-113;
 * evaluation.py:117-118.  x and y may alias.  inv_norm may be NULL. */
int trb_l2_normalize_rows_f32(const float* x, float* y, float* inv_norm, int64_t rows, int64_t dim,
                              float eps, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Retrieval evaluation  (lib/data/metrics/evaluation.py:11-37 rank(), :117-120 similarity)
 *
 * The [Q,G] similarity matrix is never written.  Work is split into
 *   thresholds : similarity of every (query, relevant gallery item) pair       (small)
 *   stream     : one pass over the gallery -> per-query top-10 candidates and, for every
 *                relevant item, the number of gallery items ranking before it   (the GEMM)
 *   finish     : merge candidate lists (gallery splits / ranks), hit ranks, AP  (small)
 *   metrics    : CMC@k and mAP scalars                                          (tiny)
 * "Relevant" = gallery pid equals the query pid (evaluation.py:20-21).  The relevant sets are
 * a CSR over queries: rel_ptr[Q+1] int64, one slot per (query, relevant item).
 * A gallery may be one shard of a larger one: g_base is the global index of local row 0 and
 * every index leaving the library is global.
 * ---------------------------------------------------------------------------------------- */

/* thr[slot] = <qn[q,:], gn[rel_row[slot],:]> accumulated in k-ascending FFMA order, bit-identical
 * to the value the fp32 stream kernel computes for the same pair.  rel_row holds LOCAL gallery rows;
 * slots whose rel_row is < 0 (item lives on another shard) are left untouched. */
int trb_retrieval_thresholds_f32(const float* qn, const float* gn, const int64_t* rel_ptr,
                                 const int64_t* rel_row, float* thr, int64_t Q, int64_t D,
                                 trb_stream_t stream);

/* One pass over a (shard of the) gallery in fp32 FFMA arithmetic.
 *   qn [Q,D], gn [G,D]   L2-normalised rows
 *   rel_ptr [Q+1], thr [total], thr_gidx [total]  thresholds and the GLOBAL gallery index of each
 *                         relevant item (tie-break); may all be NULL when ranks are not wanted
 *   nsplit               the gallery is cut into nsplit contiguous pieces processed by different CTAs
 *   cand_sim/cand_idx    [Q, nsplit, 10] best-first candidates of each piece (idx global, int64;
 *                         unused entries: -inf / INT64_MAX)
 *   cnt [total] int32    += #{local g : (s_g, g) ranks before (thr, thr_gidx)}; caller zeroes it
 */
int trb_retrieval_stream_f32(const float* qn, const float* gn, int64_t Q, int64_t G, int64_t D,
                             int64_t g_base, const int64_t* rel_ptr, const float* thr,
                             const int64_t* thr_gidx, int nsplit, float* cand_sim, int64_t* cand_idx,
                             int32_t* cnt, trb_stream_t stream);

/* Same contract as the thresholds/stream pair, from a MATERIALISED similarity matrix
 * (drop-in for rank(similarity, ...), evaluation.py:11).  sim is [Q,G] with element strides
 * (row_stride, col_stride) so that similarity.t() needs no copy.  rel_col holds gallery columns.
 * Produces one candidate list per query (nsplit = 1) and cnt (may be NULL with rel_ptr NULL). */
int trb_rank_similarity_f32(const float* sim, int64_t row_stride, int64_t col_stride, int64_t Q,
                            int64_t G, const int64_t* rel_ptr, const int64_t* rel_col,
                            float* cand_sim, int64_t* cand_idx, int32_t* cnt, trb_stream_t stream);

/* k-reciprocal re-ranking (evaluation.py:40-65, 151-156): rank the float64 scores
 *   alpha * |N(q) & N(g)| / |N(q) | N(g)|  +  sim[q,g]
 * where N(q) = q_nn[q, 0:n] are the n best gallery items of query q and N(g) = g_nn[g, 0:n] the n best gallery items
 * of gallery item g (n = 5, alpha = 0.05 in the reference).  Same outputs as trb_rank_similarity_f32; the scores are
 * formed on the fly in float64 (the reference's jaccard matrix is float64), the matrix is not materialised. */
int trb_rank_rerank_f64(const float* sim, int64_t row_stride, int64_t col_stride, int64_t Q, int64_t G,
                        const int64_t* q_nn, const int64_t* g_nn, int n_neighbors, double alpha,
                        const int64_t* rel_ptr, const int64_t* rel_col, float* cand_sim, int64_t* cand_idx,
                        int32_t* cnt, trb_stream_t stream);

/* Same as trb_rank_similarity_f32 for a materialised FLOAT64 score matrix (re-ranked scores rebuilt from the
 * rvn_mat / rtn_mat arrays of a cached inference_data.npz, evaluation.py:85-95). */
int trb_rank_scores_f64(const double* scores, int64_t row_stride, int64_t col_stride, int64_t Q, int64_t G,
                        const int64_t* rel_ptr, const int64_t* rel_col, float* cand_sim, int64_t* cand_idx,
                        int32_t* cnt, trb_stream_t stream);

/* out[Q,G] (float64) = alpha * Jaccard(N(q), N(g)): the rvn_mat / rtn_mat arrays of inference_data.npz
 * (evaluation.py:126-142).  Compatibility only. */
int trb_jaccard_f64(const int64_t* q_nn, const int64_t* g_nn, int n_neighbors, double alpha, double* out, int64_t Q,
                    int64_t G, trb_stream_t stream);

/* sim[Q,G] = qn @ gn^T, materialised (evaluation.py:120).  Compatibility only: the npz cache
 * (evaluation.py:126-142), re-ranking, and callers of rank().  Same k-ascending FFMA order as above. */
int trb_similarity_f32(const float* qn, const float* gn, float* sim, int64_t Q, int64_t G, int64_t D,
                       trb_stream_t stream);

/* Merge `nlists` candidate lists per query into the final top-10 and derive the per-query
 * ranking artefacts.
 *   cand_sim/cand_idx [Q, nlists, 10]
 *   q_pids [Q], g_pids [G_total] (global gallery pids): only read for the top-k-only first hit; may be NULL otherwise
 *   rel_ptr/cnt       NULL in top-k-only mode (rank(get_mAP=False), evaluation.py:16-19)
 * outputs
 *   top_sim/top_idx [Q,10]  best-first; idx int64 global
 *   first_hit [Q] int32     0-based rank of the best relevant item; in top-k-only mode the
 *                           position of the first pid match inside the top-10, else INT32_MAX
 *   hit_ranks [total] int32 per query, ascending 0-based ranks of its relevant items (NULL ok)
 *   ap [Q]                  sum_j fl((j+1)/(rank_j+1)) / num_rel, rank-ascending fp32 sum;
 *                           NaN when num_rel == 0 like the reference's 0/0 (NULL ok)
 */
int trb_retrieval_finish(const float* cand_sim, const int64_t* cand_idx, int nlists, int64_t Q,
                         const int64_t* q_pids, const int64_t* g_pids, int64_t G_total,
                         const int64_t* rel_ptr, const int32_t* cnt, float* top_sim, int64_t* top_idx,
                         int32_t* first_hit, int32_t* hit_ranks, float* ap, trb_stream_t stream);

/* cmc[i] = fl(#{q : first_hit[q] < topk[i]} / Q) * 100 and mAP = mean(ap) * 100
 * (evaluation.py:23-26,35-36).  The AP mean is a fixed-order fp64 reduction rounded once to fp32;
 * the bit-exact-to-torch-CPU mean is taken on the host by the Python layer in parity mode.
 * topk is a HOST array of n <= 8 cut-offs.  map_out may be NULL. */
int trb_retrieval_metrics(const int32_t* first_hit, const float* ap, int64_t Q, const int32_t* host_topk,
                          int n_topk, float* cmc_out, float* map_out, trb_stream_t stream);

/* ---- bf16 tensor-core path (tcgen05 + TMEM, operands staged by cp.async.bulk) ------------
 * Operands live in HBM in a packed, tile-major, pre-swizzled bf16 layout that is byte-identical
 * to the shared-memory image tcgen05.mma reads (128-row x 64-column K-major SWIZZLE_128B blocks),
 * so a tile load is one contiguous bulk copy.  trb_pack_rows_bf16 produces it, fusing the L2
 * normalisation (evaluation.py:117-118) and an optional row gather. */
int64_t trb_packed_rows(int64_t rows);                 /* rows rounded up to the 128-row block */
int64_t trb_packed_bytes(int64_t rows, int64_t dim);   /* bytes of the packed image, 0 if dim % 64 */

/* src: [*, dim] fp32 (src_is_bf16 = 0) or bf16 (1).  Row r of the packed image is source row
 * perm[r] (perm may be NULL = identity); rows >= `rows` are zero.  normalize != 0 applies
 * x / max(||x||, eps) in fp32 before the single rounding to bf16. */
int trb_pack_rows_bf16(const void* src, int src_is_bf16, const int64_t* perm, int normalize, float eps,
                       void* packed, int64_t rows, int64_t dim, trb_stream_t stream);

/* Candidate lists the tensor-core stream writes per query and gallery split (one per epilogue column group). */
int trb_retrieval_tc_lists_per_split(void);

/* Tensor-core pass over a packed gallery shard (G local rows; a shard holds < 2^31 rows).
 * q_row_id [Qp]: original query number of every packed query row (-1 = padding); queries may be packed in any order.
 *
 * mode 0 -- stream: the gallery MUST be packed in index order (perm = NULL); the global index of packed row g is
 *   g_base + g.  Produces top-10 candidate lists (trb_retrieval_tc_lists_per_split() * nsplit per query, best first,
 *   global int64 indices, unused lists padded with -inf / INT64_MAX) and, when rel_ptr/thr/thr_gidx/cnt are given,
 *   cnt[slot] += #{local g : (s_g, g) ranks before (thr[slot], thr_gidx[slot])}.  Because the stream is in index order,
 *   ties are resolved by position: chunks before the relevant item compare with >=, chunks after it with >.
 *   nsplit: gallery pieces for the query tiles of the last, partial wave of the persistent grid (whole waves are
 *   never split); max_rel: upper bound on the relevant items of a query (4 or 8 thresholds stay in registers, rows
 *   with more take an exact slow path).
 * mode 1 -- threshold capture: queries AND gallery packed in pid order (perm = argsort of the pids), g_row_id [Gp] gives
 *   the global index of every packed gallery row (-1 = padding).  For packed query row i the packed gallery rows
 *   [band_lo[i], band_hi[i]) are its relevant items; writes thr[rel_ptr[q] + rel_off[i] + (g - band_lo[i])] and
 *   thr_gidx likewise, with the same tcgen05.mma sequence as mode 0 so the values are bit-identical to the streamed ones. */
int trb_retrieval_stream_tc(const void* q_packed, const void* g_packed, int64_t Q, int64_t G, int64_t D,
                            const int64_t* q_row_id, const int64_t* g_row_id, int64_t g_base, const int64_t* rel_ptr,
                            float* thr, int64_t* thr_gidx, const int32_t* band_lo, const int32_t* band_hi,
                            const int32_t* rel_off, int mode, int nsplit, int max_rel, float* cand_sim,
                            int64_t* cand_idx, int32_t* cnt, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MoCo loss step  (lib/models/embeddings/moco_head/head.py:126-175, moco_head/loss.py:21-39,
 *                  lib/models/losses.py:6-62,102-128,206-217)
 * ---------------------------------------------------------------------------------------- */

typedef struct trb_moco_shape {
    int32_t N;      /* batch */
    int32_t D;      /* embedding size, cfg.MODEL.EMBEDDING.FEATURE_SIZE */
    int32_t K;      /* queue length, cfg.MODEL.MOCO.K */
    int32_t C;      /* cfg.MODEL.NUM_CLASSES */
} trb_moco_shape;

typedef struct trb_moco_hparams {
    float T;            /* 0.07, moco_head/loss.py:18 */
    float epsilon;      /* label smoothing, cfg.MODEL.EMBEDDING.EPSILON */
    float alpha, beta;  /* 0.6, 0.4  losses.py:106-107 */
    float scale_pos, scale_neg; /* 10, 40 losses.py:108-109 */
} trb_moco_hparams;

int64_t trb_moco_loss_workspace_bytes(const trb_moco_shape* shape, int precision);

/* Kernel launches one trb_moco_loss call issues for this shape on the current device: 2 when precision = 1 takes the fused
 * path (a small prologue + ONE cooperative tcgen05 kernel for the three losses and all gradients; needs D a multiple of 64 up
 * to 256 and, for N <= 128, ceil(C/128) + 2*ceil(K/128) + 1 <= #SMs or, for 128 < N <= 256, ceil(C/128) + ceil(K/128) + 2 <=
 * #SMs), otherwise the length of the launch sequence (N up to 1024: instance and InfoNCE branches still one cooperative launch).
 * The cooperative kernel occupies every SM of the device (one CTA per class / queue tile, spare CTAs on the rest that share the
 * final reductions) and is launched as a programmatic dependent launch behind the prologue (it starts while the prologue still
 * runs and waits for it on the device); both survive CUDA-graph capture.  Batches above 128 rows are walked in 128-row windows
 * inside the kernel. */
int trb_moco_loss_launches(const trb_moco_shape* shape, int precision);

/* Loss dict and its gradients in one stream-ordered call (fwd and bwd fused: the softmax
 * statistics are consumed where they are produced, nothing is saved for a later backward).
 *   v_embed, t_embed [N,D]   post-Linear, un-normalised (instance + global-align inputs)
 *   v_qraw, t_qraw   [N,D]   InfoNCE query inputs before normalisation; pass the embeds
 *                            themselves when cfg.MODEL.MOCO.FC is False (head.py:126-129)
 *   v_key, t_key     [N,D]   key embeddings; normalised here when normalize_keys != 0
 *                            (head.py:139,145) and written to v_key_n / t_key_n [N,D]
 *   labels [N] int64; v_queue, t_queue [D,K]; id_queue [K] int64 (-1 = empty slot)
 *   projection [D,C]
 *   precision: 0 = fp32 FFMA path (parity, 1e-5); 1 = bf16 tcgen05 path (1e-3): operands rounded once to bf16, fp32
 *              accumulation in TMEM, softmax / loss math in fp32.  Shapes that fit (trb_moco_loss_launches == 2) run as ONE
 *              cooperative kernel in which logits, softmax statistics and logit gradients never leave the SM.
 * outputs
 *   losses [3]               instance, infonce, global_align (each with upstream grad 1)
 *   d_inst, d_nce, d_ga      [2,N,D] per-loss gradients w.r.t. (v,t) embeds / qraw; NULL skips bwd
 *   d_projection [D,C]       gradient of instance_loss w.r.t. projection
 * The queue column mask (head.py:148-157) is evaluated on the device; no host sync.
 * The three losses are independent between the shared prologue and the final reduction: the call forks two
 * internal helper streams off `stream` (created once per device, joined back before the call's last launch), so the
 * work is ordered after / before everything else on `stream` as usual and a CUDA-graph capture of the call yields a
 * 3-wide DAG (unfused launch sequences only; the fused bf16 path is two launches on `stream`).  The helper streams are the one
 * piece of process-level state of the library: one trb_moco_loss / trb_moco_step call of an UNFUSED shape may be in flight per
 * device at a time.
 * A label outside [0, C) (the reference raises in scatter_, losses.py:33) yields a NaN instance loss instead of an
 * out-of-bounds access. */
int trb_moco_loss(const float* v_embed, const float* t_embed, const float* v_qraw, const float* t_qraw,
                  const float* v_key, const float* t_key, int normalize_keys, float* v_key_n,
                  float* t_key_n, const int64_t* labels, const float* v_queue, const float* t_queue,
                  const int64_t* id_queue, const float* projection, const trb_moco_shape* shape,
                  const trb_moco_hparams* hp, int precision, float* losses, float* d_inst, float* d_nce,
                  float* d_ga, float* d_projection, void* workspace, int64_t workspace_bytes,
                  trb_stream_t stream);

/* trb_moco_loss followed by _dequeue_and_enqueue (head.py:175, :96-109) as ONE stream-ordered call -- what MoCoHead.forward's
 * train branch does after the encoders.  Same arguments as trb_moco_loss; the queues, id_queue and queue_ptr [1] (int64, read
 * and advanced on the device: no int(queue_ptr) sync) are written after the last read of the old queue contents.  Requires
 * K % N == 0 like the reference's assert (head.py:101).  On the fused bf16 path the enqueue rides inside the cooperative
 * kernel (the InfoNCE CTAs, idle at that point, write the key columns; the pointer moves after every tile has arrived at the kernel's last grid-wide counter), so the
 * whole step stays at trb_moco_loss_launches() launches; otherwise two small launches follow the loss sequence.
 * A queue_ptr outside [0, K-N] (a checkpoint written with another batch size) wraps modulo K instead of writing out of bounds. */
int trb_moco_step(const float* v_embed, const float* t_embed, const float* v_qraw, const float* t_qraw,
                  const float* v_key, const float* t_key, int normalize_keys, float* v_key_n, float* t_key_n,
                  const int64_t* labels, float* v_queue, float* t_queue, int64_t* id_queue, int64_t* queue_ptr,
                  const float* projection, const trb_moco_shape* shape, const trb_moco_hparams* hp, int precision,
                  float* losses, float* d_inst, float* d_nce, float* d_ga, float* d_projection, void* workspace,
                  int64_t workspace_bytes, trb_stream_t stream);

/* Kernel launches of one trb_moco_step call (2 on the fused bf16 path). */
int trb_moco_step_launches(const trb_moco_shape* shape, int precision);

/* Debug read-backs of the fused bf16 path (tests and tools only; both SYNCHRONISE and copy to HOST memory).  The data lives in
 * the caller's workspace, the library owns no memory.
 *   stamps: [160][16] uint64 %globaltimer phase stamps per CTA of the last launch made with the environment variable
 *           TRB_FUSED_DEBUG set;
 *   logits: [256][128] fp32 -- rows = (modality, batch row), columns = the 128 classes of instance tile t -- the logits
 *           z = e . W/||W||_col exactly as the kernel's softmax saw them, for the last launch made with TRB_FUSED_DEBUG_LOGITS=t.
 *           On the product path the logits never leave the SM. */
int trb_moco_loss_debug_stamps(const void* workspace, const trb_moco_shape* shape, unsigned long long* host_out);
int trb_moco_loss_debug_logits(const void* workspace, const trb_moco_shape* shape, float* host_out);

/* out = g[0]*a + g[1]*b + g[2]*c with g a DEVICE array of 3 upstream gradients (any of a,b,c may
 * be NULL).  Backward of the loss dict without a host sync (trainer.py:82,90 uses g = 1,1,1). */
int trb_combine3_f32(float* out, const float* a, const float* b, const float* c, const float* g,
                     int64_t n, trb_stream_t stream);

/* Whole backward of the loss dict in ONE launch: the per-loss gradients saved by trb_moco_loss times the three upstream
 * gradients (DEVICE scalars; NULL = that loss did not take part in the backward pass, trainer.py:82,90 passes 1,1,1):
 *   out_v / out_t [N,D]   = g_inst * d_inst[m] + g_nce * d_nce[m] + g_ga * d_ga[m]          (m = 0 image, 1 text)
 *   separate_q (cfg.MODEL.MOCO.FC, head.py:118-124): the InfoNCE term goes to out_vq / out_tq instead
 *   out_proj [D,C]        = g_inst * d_proj   (NULL skips it; may alias d_proj: then it is scaled in place, and left untouched
 *                           without any memory traffic when g_inst == 1)
 * nd = N*D, dc = D*C. */
int trb_moco_grad_combine(const float* d_inst, const float* d_nce, const float* d_ga, const float* d_proj,
                          const float* g_inst, const float* g_nce, const float* g_ga, int separate_q, int64_t nd,
                          int64_t dc, float* out_v, float* out_t, float* out_vq, float* out_tq, float* out_proj,
                          trb_stream_t stream);

/* x *= g[0] unless g[0] == 1 (then no memory traffic). */
int trb_scale_inplace_f32(float* x, const float* g, int64_t n, trb_stream_t stream);

/* Momentum update p_k <- fl(fl(p_k*m) + fl(p_q*(1-m))), exactly the reference's two products and
 * one add (head.py:78-94); one_minus_m is passed separately because the reference forms 1-m in
 * double.  Flat form: one contiguous parameter arena. */
int trb_ema_update_f32(float* p_k, const float* p_q, int64_t n, float m, float one_minus_m,
                       trb_stream_t stream);

/* Multi-tensor form: a device table of `nchunks` (k_ptr, q_ptr, count) chunks built once by the host. */
typedef struct trb_ema_chunk {
    float* k;
    const float* q;
    int64_t n;
} trb_ema_chunk;
int trb_ema_update_chunks_f32(const trb_ema_chunk* chunks, int64_t nchunks, int64_t max_chunk, float m,
                              float one_minus_m, trb_stream_t stream);

/* _dequeue_and_enqueue (head.py:96-109): queue[:, ptr:ptr+N] = keys^T for both queues and the ids,
 * then ptr = (ptr+N) % K, with ptr read and written on the device.  Requires K % N == 0 like the
 * reference's assert.  v_keys, t_keys [N,D] normalised; queues [D,K]; id_queue [K]; queue_ptr [1].
 * A pointer outside [0, K-N] wraps modulo K (the reference's slice assignment raises there). */
int trb_enqueue(float* v_queue, float* t_queue, int64_t* id_queue, int64_t* queue_ptr,
                const float* v_keys, const float* t_keys, const int64_t* ids, int32_t N, int32_t D,
                int32_t K, trb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TEXTREID_B200_H */
