// bf16 tensor-core MoCo loss (precision = 1).  Two forms, both entered through trb_moco_loss_tc (loss_f32.cu):
//   * fused (loss_fused.cu): a small prologue + ONE cooperative tcgen05 kernel for the three losses and all gradients, taken
//     whenever the shape fits (N <= 128, D % 64 == 0, D <= 256, one CTA per 128-class / 128-slot tile fits the device);
//   * unfused: the launch sequence of loss_f32.cu with every contraction going through tc_gemm.cu (operands rounded once to bf16
//     in the packed tile-major layout, tcgen05.mma with fp32 accumulation in TMEM); the row-wise softmax / loss / gradient
//     kernels stay fp32.
