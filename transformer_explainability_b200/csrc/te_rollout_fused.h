// Fused aggregation + rollout, row-only mode:
//   r <- e_0 ;  for l = L-1 .. start:  r <- r (M_l + I)     with  M_l = mean_h relu(G_l * cam_l)  (rows /rowsum if normalize)
// which is row 0 of J = (M_{L-1}+I) ... (M_start+I) — all that generate_LRP consumes (ViT_LRP.py:368,
// ExplanationGenerator.py:58-59).  One thread-block cluster per sample streams G and cam exactly once; nothing
// else touches HBM.  The dense joint (compute_rollout_attention's public result) stays on the composed path.
#pragma once
#include "te_common.cuh"
bool te_rollout_fused_supported(int N, int ld_in, int ld);
// row_out [B, N-first]; bert_fix: element 0 replaced by the row minimum
int te_rollout_fused_row(const float* G0, const float* cam0, long long layer_stride, int L, int B, int H, int N,
                         int ld_in, int start_layer, int normalize, float* row_out, int first, int bert_fix,
                         cudaStream_t st);
