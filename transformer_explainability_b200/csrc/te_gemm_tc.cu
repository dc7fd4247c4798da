// placeholder until the tcgen05 kernel lands: reports "unsupported" so callers use the fp32 SIMT path.
#include "te_gemm_tc.h"
bool te_tc_zplus_supported(long long, int, int, long long) { return false; }
int te_tc_zplus_linear_relprop(const float*, long long, const float*, const float*, float*, float*, long long, int,
                               int, cudaStream_t) {
    te_set_last_error("tcgen05 z+ path not built");
    return TE_ERR_UNSUPPORTED;
}
