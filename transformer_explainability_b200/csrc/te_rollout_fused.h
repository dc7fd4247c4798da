// Fused aggregation + rollout: grad*cam -> relu -> head-mean -> +I -> (row-normalise) -> running product,
// one persistent CTA group per sample, running product on-chip.
#pragma once
#include "te_common.cuh"
bool te_rollout_fused_supported(int N, int ld_in, int ld);
int te_rollout_fused(const float* G0, const float* cam0, long long layer_stride, int L, int B, int H, int N, int ld_in,
                     int ld, int start_layer, int normalize, float* joint /*[B,N,ld]*/, cudaStream_t st);
