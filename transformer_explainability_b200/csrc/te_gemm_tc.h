// tcgen05 (5th-gen tensor core) path of the z+ Linear rule — TF32 operands, fp32 accumulation in TMEM.
#pragma once
#include "te_common.cuh"

// shapes the tensor-core path accepts (in/out multiples of 256, 16-byte aligned rows)
bool te_tc_zplus_supported(long long rows, int in_features, int out_features, long long ldx);
// derived copies of one frozen weight W [out,in]: W+, W- (K-major for S-kernel), W+^T, W-^T (K-major for R-kernel),
// rounded to TF32: 4*in*out floats
long long te_tc_derived_floats(int in_features, int out_features);
int te_tc_prepare_weights(const float* w, float* derived, int in_features, int out_features, cudaStream_t st);
int te_tc_zplus_linear_relprop(const float* x, long long ldx, const float* derived, const float* r, long long ldr,
                               float* out,
                               float* s_scratch, long long rows, int in_features, int out_features, cudaStream_t st);
