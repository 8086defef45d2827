// The bf16 tensor-core MoCo loss path (precision = 1) shares the launch sequence of loss_f32.cu; every contraction goes
// through tc_gemm.cu (operands rounded once to bf16 in the packed tile-major layout, tcgen05.mma with fp32 accumulation in
// TMEM).  The row-wise softmax / loss / gradient kernels stay fp32.  Entry points: trb_moco_loss_tc,
// trb_moco_loss_workspace_bytes_tc (defined in loss_f32.cu next to the shared implementation).
