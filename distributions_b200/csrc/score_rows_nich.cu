// score_rows_nich.cu -- single-feature instantiations of score_rows_kernel for one model (see score_rows.cuh)
#include "score_rows.cuh"

namespace distb200 {

int launch_single_nich(dist_b200_ctx *ctx, const FeatList &feats, const RowsArgs &a, cudaStream_t s) {
    // default: the packed fp32x2 loop; DIST_B200_OPT_NICH_PACKED = 1 keeps the scalar loop for A/B runs
    if (ctx->opt[DIST_B200_OPT_NICH_PACKED] == 1) return launch_tiers<DIST_B200_NICH>(ctx, feats, a, s);
    return launch_tiers<kKindNichPacked>(ctx, feats, a, s);
}

}  // namespace distb200
