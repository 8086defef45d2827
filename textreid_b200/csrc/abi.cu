// C-ABI plumbing: version, per-thread error string, precision dispatch of the loss entry point.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void trb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int trb_version(void) { return TRB_VERSION; }
extern "C" const char* trb_last_error_string(void) { return g_err; }

int64_t trb_moco_loss_workspace_bytes_f32(const trb_moco_shape* s);
int trb_moco_loss_f32(const float*, const float*, const float*, const float*, const float*, const float*, int, float*, float*,
                      const int64_t*, const float*, const float*, const int64_t*, const float*, const trb_moco_shape*,
                      const trb_moco_hparams*, float*, float*, float*, float*, float*, void*, int64_t, cudaStream_t, int64_t*);
int64_t trb_moco_loss_workspace_bytes_tc(const trb_moco_shape* s);
int trb_moco_loss_tc(const float*, const float*, const float*, const float*, const float*, const float*, int, float*, float*,
                     const int64_t*, const float*, const float*, const int64_t*, const float*, const trb_moco_shape*,
                     const trb_moco_hparams*, float*, float*, float*, float*, float*, void*, int64_t, cudaStream_t, int64_t*);
int trb_moco_step_extra_launches_impl(const trb_moco_shape* s, int precision);
int trb_moco_loss_debug_impl(const void* workspace, const trb_moco_shape* s, int what, void* host_out);

static int check_shape(const trb_moco_shape* s) {
    TRB_REQUIRE(s != nullptr, "moco_loss: null shape");
    TRB_REQUIRE(s->N > 0 && s->D > 0 && s->K > 0 && s->C > 0, "moco_loss: non-positive dimension N=%d D=%d K=%d C=%d", s->N,
                s->D, s->K, s->C);
    return 0;
}

extern "C" int64_t trb_moco_loss_workspace_bytes(const trb_moco_shape* shape, int precision) {
    if (check_shape(shape)) return TRB_ERR_INVALID;
    if (precision == 0) return trb_moco_loss_workspace_bytes_f32(shape);
    if (precision == 1) return trb_moco_loss_workspace_bytes_tc(shape);
    trb_set_error("moco_loss: unknown precision %d", precision);
    return TRB_ERR_INVALID;
}

int trb_moco_loss_launches_impl(const trb_moco_shape* s, int precision);
extern "C" int trb_moco_loss_launches(const trb_moco_shape* shape, int precision) {
    if (check_shape(shape)) return TRB_ERR_INVALID;
    return trb_moco_loss_launches_impl(shape, precision);
}

static int moco_step_checked(const float* v_embed, const float* t_embed, const float* v_qraw, const float* t_qraw,
                             const float* v_key, const float* t_key, int normalize_keys, float* v_key_n, float* t_key_n,
                             const int64_t* labels, const float* v_queue, const float* t_queue, const int64_t* id_queue,
                             int64_t* queue_ptr, const float* projection, const trb_moco_shape* shape, const trb_moco_hparams* hp,
                             int precision, float* losses, float* d_inst, float* d_nce, float* d_ga, float* d_projection,
                             void* workspace, int64_t workspace_bytes, trb_stream_t stream) {
    int rc = check_shape(shape);
    if (rc) return rc;
    TRB_REQUIRE(hp != nullptr, "moco_loss: null hyper-parameters");
    TRB_REQUIRE(v_embed && t_embed && v_qraw && t_qraw && v_key && t_key && v_key_n && t_key_n && labels && v_queue &&
                    t_queue && id_queue && projection && losses && workspace,
                "moco_loss: null pointer");
    TRB_REQUIRE((d_inst == nullptr) == (d_nce == nullptr) && (d_inst == nullptr) == (d_ga == nullptr),
                "moco_loss: d_inst, d_nce and d_ga must be given together");
    TRB_REQUIRE(d_projection == nullptr || d_inst != nullptr, "moco_loss: d_projection needs the embedding gradients too");
    TRB_REQUIRE(hp->T > 0.f, "moco_loss: temperature must be positive");
    if (queue_ptr != nullptr)
        TRB_REQUIRE(shape->K % shape->N == 0, "moco_step: K=%d must be a multiple of the batch size N=%d (head.py:101)", shape->K,
                    shape->N);
    if (precision == 0)
        return trb_moco_loss_f32(v_embed, t_embed, v_qraw, t_qraw, v_key, t_key, normalize_keys, v_key_n, t_key_n, labels,
                                 v_queue, t_queue, id_queue, projection, shape, hp, losses, d_inst, d_nce, d_ga, d_projection,
                                 workspace, workspace_bytes, (cudaStream_t)stream, queue_ptr);
    if (precision == 1)
        return trb_moco_loss_tc(v_embed, t_embed, v_qraw, t_qraw, v_key, t_key, normalize_keys, v_key_n, t_key_n, labels,
                                v_queue, t_queue, id_queue, projection, shape, hp, losses, d_inst, d_nce, d_ga, d_projection,
                                workspace, workspace_bytes, (cudaStream_t)stream, queue_ptr);
    trb_set_error("moco_loss: unknown precision %d", precision);
    return TRB_ERR_INVALID;
}

extern "C" int trb_moco_loss(const float* v_embed, const float* t_embed, const float* v_qraw, const float* t_qraw,
                             const float* v_key, const float* t_key, int normalize_keys, float* v_key_n, float* t_key_n,
                             const int64_t* labels, const float* v_queue, const float* t_queue, const int64_t* id_queue,
                             const float* projection, const trb_moco_shape* shape, const trb_moco_hparams* hp, int precision,
                             float* losses, float* d_inst, float* d_nce, float* d_ga, float* d_projection, void* workspace,
                             int64_t workspace_bytes, trb_stream_t stream) {
    return moco_step_checked(v_embed, t_embed, v_qraw, t_qraw, v_key, t_key, normalize_keys, v_key_n, t_key_n, labels, v_queue,
                             t_queue, id_queue, nullptr, projection, shape, hp, precision, losses, d_inst, d_nce, d_ga,
                             d_projection, workspace, workspace_bytes, stream);
}

extern "C" int trb_moco_step(const float* v_embed, const float* t_embed, const float* v_qraw, const float* t_qraw,
                             const float* v_key, const float* t_key, int normalize_keys, float* v_key_n, float* t_key_n,
                             const int64_t* labels, float* v_queue, float* t_queue, int64_t* id_queue, int64_t* queue_ptr,
                             const float* projection, const trb_moco_shape* shape, const trb_moco_hparams* hp, int precision,
                             float* losses, float* d_inst, float* d_nce, float* d_ga, float* d_projection, void* workspace,
                             int64_t workspace_bytes, trb_stream_t stream) {
    TRB_REQUIRE(queue_ptr != nullptr, "moco_step: null queue pointer");
    return moco_step_checked(v_embed, t_embed, v_qraw, t_qraw, v_key, t_key, normalize_keys, v_key_n, t_key_n, labels, v_queue,
                             t_queue, id_queue, queue_ptr, projection, shape, hp, precision, losses, d_inst, d_nce, d_ga,
                             d_projection, workspace, workspace_bytes, stream);
}

extern "C" int trb_moco_step_launches(const trb_moco_shape* shape, int precision) {
    if (check_shape(shape)) return TRB_ERR_INVALID;
    return trb_moco_loss_launches_impl(shape, precision) + trb_moco_step_extra_launches_impl(shape, precision);
}

extern "C" int trb_moco_loss_debug_stamps(const void* workspace, const trb_moco_shape* shape, unsigned long long* host_out) {
    if (check_shape(shape)) return TRB_ERR_INVALID;
    TRB_REQUIRE(workspace && host_out, "moco_loss debug: null pointer");
    return trb_moco_loss_debug_impl(workspace, shape, 0, host_out);
}

extern "C" int trb_moco_loss_debug_logits(const void* workspace, const trb_moco_shape* shape, float* host_out) {
    if (check_shape(shape)) return TRB_ERR_INVALID;
    TRB_REQUIRE(workspace && host_out, "moco_loss debug: null pointer");
    return trb_moco_loss_debug_impl(workspace, shape, 1, host_out);
}
