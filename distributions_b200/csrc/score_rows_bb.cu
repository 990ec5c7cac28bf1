// score_rows_bb.cu -- single-feature instantiations of score_rows_kernel for one model (see score_rows.cuh)
#include "score_rows.cuh"

namespace distb200 {

int launch_single_bb(dist_b200_ctx *ctx, const FeatList &feats, const RowsArgs &a, cudaStream_t s) {
    return launch_tiers<DIST_B200_BB>(ctx, feats, a, s);
}

}  // namespace distb200
