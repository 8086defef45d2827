// bf16 tcgen05 MoCo loss path (precision = 1).  Placeholder until the tensor-core tile core of
// retrieval_tc.cu is reused here: the entry points reject the request loudly, never fall back.
#include "common.cuh"

int64_t trb_moco_loss_workspace_bytes_tc(const trb_moco_shape*) {
    trb_set_error("moco_loss: the bf16 tensor-core path is not built yet");
    return TRB_ERR_UNSUPPORTED;
}

int trb_moco_loss_tc(const float*, const float*, const float*, const float*, const float*, const float*, int, float*, float*,
                     const int64_t*, const float*, const float*, const int64_t*, const float*, const trb_moco_shape*,
                     const trb_moco_hparams*, float*, float*, float*, float*, float*, void*, int64_t, cudaStream_t) {
    trb_set_error("moco_loss: the bf16 tensor-core path is not built yet");
    return TRB_ERR_UNSUPPORTED;
}
