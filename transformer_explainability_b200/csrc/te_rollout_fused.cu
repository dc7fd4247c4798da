// placeholder until the fused kernel lands
#include "te_rollout_fused.h"
bool te_rollout_fused_supported(int, int, int) { return false; }
int te_rollout_fused(const float*, const float*, long long, int, int, int, int, int, int, int, int, float*, cudaStream_t) {
    te_set_last_error("fused rollout not built");
    return TE_ERR_UNSUPPORTED;
}
