// Fused bf16 MoCo loss step (loss_fused.cu): one prologue launch + ONE cooperative tcgen05 kernel for the whole loss dict and
// its gradients.  Used by trb_moco_loss(precision = 1) when the shape fits (see fused_loss_supported).
#pragma once
#include "common.cuh"

struct FusedLossArgs {
    int N, D, K, C;
    float T, eps, alpha, beta, sp, sn;
    // inputs
    const float *v_embed, *t_embed, *v_qraw, *t_qraw, *v_key, *t_key;
    int normalize_keys;
    const int64_t *labels, *id_queue;
    const float *v_queue, *t_queue, *projection;
    // fp32 row artefacts in the caller's workspace (shared with the unfused launch sequence)
    float *v_key_n, *t_key_n, *E2, *en, *qn, *inv_e, *inv_q, *pos, *dpos, *rows_inst, *rows_nce, *rows_ga;
    // outputs
    float *losses, *d_inst, *d_nce, *d_ga, *d_proj;
    // fused scratch (fused_loss_scratch_bytes), 1024-byte aligned
    uint8_t* scratch;
    // _dequeue_and_enqueue folded into the cooperative kernel (all roles fused only); NULL enq_ptr = no enqueue
    float *enq_v_queue, *enq_t_queue;
    int64_t *enq_ids, *enq_ptr;
    int roles;          // bit 0 instance, bit 1 InfoNCE, bit 2 global-align: which branches the fused kernel runs
    int reduce_losses;  // the fused kernel also forms the three loss scalars (all roles fused)
    int after_prologue; // this launch directly follows fused_loss_prologue on the stream (programmatic dependent launch allowed)
};

// shape gate: D a multiple of 64 up to 256, N <= 128, and one CTA per 128-class / 128-slot tile must fit the device
bool fused_loss_supported(int N, int D, int K, int C, int sm_count);
// 128 < N <= 1024: the kernel walks the batch in row windows of 128 (W / queue tiles resident); an InfoNCE CTA takes its tile
// index for both modalities in turn, so ceil(C/128) + ceil(K/128) CTAs must fit the device (86 + 32 at C = 11003, K = 4096)
bool fused_windows_supported(int N, int D, int K, int C, int sm_count);
int64_t fused_loss_scratch_bytes(int N, int D, int K, int C);
// prologue (row norms, positive logits, packed bf16 operand images, barrier reset) on `st`
int fused_loss_prologue(const FusedLossArgs& a, cudaStream_t st);
// the cooperative kernel on `st` (after the prologue)
int fused_loss_launch(const FusedLossArgs& a, cudaStream_t st);
// debug: copy the phase timestamps (what = 0, [160][16] u64) or the dumped logits tile (what = 1, [256][128] fp32) of the last
// fused launch on this scratch to the host (synchronous)
int fused_loss_debug_copy(const uint8_t* scratch, int N, int D, int K, int C, int what, void* host_out);
