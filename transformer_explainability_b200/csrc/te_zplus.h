// z+ rule of Linear.relprop (modules/layers_ours.py:207-230, alpha=1):
//   Z = x+ W+^T + x- W-^T ; S = safe_divide(R, Z) ; R_in = x+ * (S W+) + x- * (S W-)
#pragma once
#include "te_common.cuh"

// x [rows, in] with row stride ldx ; w [out, in] ; r [rows, out] ; out [rows, in] ; s_scratch [rows, out].
// w_derived: the te_tc_prepare_weights() copies of w, or NULL.  When given (and the shape qualifies) both
// contractions run on tcgen05 tensor cores (TF32 inputs, fp32 accumulate); otherwise — and as the checker —
// the fp32 SIMT path.
int te_zplus_linear_relprop(const float* x, long long ldx, const float* w, const float* w_derived, const float* r,
                            float* out, float* s_scratch, long long rows, int in_features, int out_features,
                            cudaStream_t st);
// same with an explicit row stride for r (a column slice of a packed [rows, 3*out] relevance tensor)
// y / bias: the Linear's saved forward output and bias (optional; enables the single-pass tensor-core S kernel)
int te_zplus_linear_relprop_ldr(const float* x, long long ldx, const float* w, const float* w_derived, const float* r,
                                long long ldr, float* out, float* s_scratch, long long rows, int in_features,
                                int out_features, cudaStream_t st, const float* y = nullptr, long long ldy = 0,
                                const float* bias = nullptr, int bf16 = 0, long long ld_out = 0,
                                float* xabs = nullptr);
// bf16: bit 0 round-1 bf16 R kernel (TE_FLAG_ZPLUS_BF16), bit 1 bf16 |x||W|^T term (TE_FLAG_ZPLUS_S1_BF16), bit 2 second
// contraction on tcgen05 kind::f16 with a block-scaled fp16 S written by the S kernel's epilogue (TE_FLAG_ZPLUS_R_F16)
// xabs: scratch [rows, in] (the tf32(|x|) operand of the persistent single-pass S kernel); without it the tensor-core path
// uses the round-1 kernels.
// ld_out: row stride of out (0 = in_features).  With row strides on x, r, y and out the rule runs on a strided subset of
// token rows — the CLS rows of the top block, the only rows whose relevance is non-zero there.

// Linear.relprop of the layers_lrp baseline variant (modules/layers_lrp.py:187-210, alpha=1): S1 = sd(R, x+ W+^T),
// S2 = sd(R, x- W-^T) (separate denominators), R_in = x+ * (S1 W+) + x- * (S2 W-).  fp32 SIMT; s_scratch [rows, out].
int te_zplus_linear_relprop_lrp(const float* x, long long ldx, const float* w, const float* r, long long ldr, float* out,
                                float* s_scratch, long long rows, int in_features, int out_features, cudaStream_t st);
