// placeholder until the registration chain lands
#include "sb_common.cuh"
int sb_register_pairs_impl(sb_ctx* ctx, const sb_register_job*, sb_pair_result*) {
    return sb_fail(ctx, SB_ERR_UNSUPPORTED, "registration not built yet");
}
int sb_normalize_impl(sb_ctx* ctx, const void*, void*, int, int, int, int, int) {
    return sb_fail(ctx, SB_ERR_UNSUPPORTED, "normalize not built yet");
}
