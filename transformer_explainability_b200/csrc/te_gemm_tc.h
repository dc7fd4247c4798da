// tcgen05 (5th-gen tensor core) path of the z+ Linear rule — TF32 operands, fp32 accumulation in TMEM.
#pragma once
#include "te_common.cuh"

bool te_tc_zplus_supported(long long rows, int in_features, int out_features, long long ldx);
int te_tc_zplus_linear_relprop(const float* x, long long ldx, const float* w, const float* r, float* out,
                               float* s_scratch, long long rows, int in_features, int out_features, cudaStream_t st);
