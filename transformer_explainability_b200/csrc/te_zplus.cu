// z+ Linear rule driver: fp32 SIMT path (te_gemm.cu) or tcgen05 tensor-core path (te_tc_zplus.cu).
#include "te_zplus.h"
#include "te_gemm.cuh"
#include "te_gemm_tc.h"
#include <string.h>

int te_zplus_linear_relprop(const float* x, long long ldx, const float* w, const float* w_derived, const float* r,
                            float* out, float* s_scratch, long long rows, int in_features, int out_features,
                            cudaStream_t st) {
    return te_zplus_linear_relprop_ldr(x, ldx, w, w_derived, r, out_features, out, s_scratch, rows, in_features,
                                       out_features, st);
}

int te_zplus_linear_relprop_ldr(const float* x, long long ldx, const float* w, const float* w_derived, const float* r,
                                long long ldr, float* out, float* s_scratch, long long rows, int in_features,
                                int out_features, cudaStream_t st, const float* y, long long ldy, const float* bias, int bf16,
                                long long ld_out, float* xabs) {
    if (rows <= 0) return TE_OK;
    if (ld_out == 0) ld_out = in_features;
    if (rows > 0x7fffffffLL || ldx > 0x7fffffffLL) { te_set_last_error("zplus: rows/ldx overflow int"); return TE_ERR_ARG; }
    if (w_derived && ldr % 4 == 0 && ld_out % 4 == 0 && te_tc_zplus_supported(rows, in_features, out_features, ldx)) {
        const int rc = te_tc_zplus_linear_relprop(x, ldx, w_derived, r, ldr, out, s_scratch, rows, in_features, out_features, st, y,
                                                  ldy, bias, bf16, ld_out, xabs);
        if (rc != TE_ERR_UNSUPPORTED) return rc;
    }
    if (ld_out > 0x7fffffffLL) { te_set_last_error("zplus: ld_out overflow int"); return TE_ERR_ARG; }
    TeGemm p;
    memset(&p, 0, sizeof(p));
    p.nb1 = p.nb2 = 1; p.alpha = 1.f;
    // S = sd(R, x+ W+^T + x- W-^T)
    p.A = x; p.lda = (int)ldx; p.B = w; p.ldb = in_features; p.C = s_scratch; p.ldc = out_features;
    p.E0 = r; p.lde0 = (int)ldr; p.M = (int)rows; p.N = out_features; p.K = in_features;
    TE_TRY(te_gemm_launch(p, TE_L_K, TE_L_K, TE_XF_AB_POSNEG, TE_EPI_SD, st));
    // R_in = x+ * (S W+) + x- * (S W-)
    p.A = s_scratch; p.lda = out_features; p.B = w; p.ldb = in_features; p.C = out; p.ldc = (int)ld_out;
    p.E0 = x; p.lde0 = (int)ldx; p.M = (int)rows; p.N = in_features; p.K = out_features;
    TE_TRY(te_gemm_launch(p, TE_L_K, TE_L_MN, TE_XF_B_POS, TE_EPI_MULPOS, st));
    TE_TRY(te_gemm_launch(p, TE_L_K, TE_L_MN, TE_XF_B_NEG, TE_EPI_MULNEG_ACC, st));
    return TE_OK;
}

int te_zplus_linear_relprop_lrp(const float* x, long long ldx, const float* w, const float* r, long long ldr, float* out,
                                float* s_scratch, long long rows, int in_features, int out_features, cudaStream_t st) {
    if (rows <= 0) return TE_OK;
    if (rows > 0x7fffffffLL || ldx > 0x7fffffffLL || ldr > 0x7fffffffLL) { te_set_last_error("zplus_lrp: overflow"); return TE_ERR_ARG; }
    TeGemm p;
    memset(&p, 0, sizeof(p));
    p.nb1 = p.nb2 = 1; p.alpha = 1.f;
    for (int half = 0; half < 2; ++half) {
        // S_half = sd(R, x+- W+-^T)
        p.A = x; p.lda = (int)ldx; p.B = w; p.ldb = in_features; p.C = s_scratch; p.ldc = out_features;
        p.E0 = r; p.lde0 = (int)ldr; p.M = (int)rows; p.N = out_features; p.K = in_features;
        TE_TRY(te_gemm_launch(p, TE_L_K, TE_L_K, half ? TE_XF_AB_NEG : TE_XF_AB_POS, TE_EPI_SD, st));
        // R_in (+)= x+- * (S_half W+-)
        p.A = s_scratch; p.lda = out_features; p.B = w; p.ldb = in_features; p.C = out; p.ldc = in_features;
        p.E0 = x; p.lde0 = (int)ldx; p.M = (int)rows; p.N = in_features; p.K = out_features;
        TE_TRY(te_gemm_launch(p, TE_L_K, TE_L_MN, half ? TE_XF_B_NEG : TE_XF_B_POS, half ? TE_EPI_MULNEG_ACC : TE_EPI_MULPOS, st));
    }
    return TE_OK;
}
